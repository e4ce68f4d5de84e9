set -x
timeout -k 5 600 python -m pytest tests -m gpu -q -x --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_gemv2.log 2>&1; tail -15 gpurun_out/test_gemv2.log
timeout 300 python scripts/bench_linear.py 1 2 4 > gpurun_out/bench_linear_gemv2.log 2>&1; cat gpurun_out/bench_linear_gemv2.log
timeout -k 5 600 python bench.py --no-cpu-baseline > gpurun_out/bench_gemv2.log 2>&1; tail -c 1800 gpurun_out/bench_gemv2.log
