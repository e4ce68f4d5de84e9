set -x
timeout -k 5 600 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_quick.log 2>&1; tail -15 gpurun_out/test_quick.log
timeout -k 5 600 python bench.py --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; tail -c 1800 gpurun_out/bench_quick.log
