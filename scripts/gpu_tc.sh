set -x
timeout -k 5 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "tensor_core" --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_tc.log 2>&1; tail -30 gpurun_out/test_tc.log
