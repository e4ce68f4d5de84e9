set -x
timeout -k 5 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "tensor_core" --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_tc.log 2>&1; tail -5 gpurun_out/test_tc.log
timeout 300 python scripts/bench_linear.py > gpurun_out/bench_linear.log 2>&1; cat gpurun_out/bench_linear.log
B200_FORCE_TC=1 timeout 300 python scripts/bench_linear.py 1 4 > gpurun_out/bench_linear_tc.log 2>&1; cat gpurun_out/bench_linear_tc.log
