set -x
timeout -k 5 300 python -m pytest tests/test_weights.py -x -q -m gpu --timeout 200 -p no:cacheprovider > gpurun_out/w_tests.log 2>&1; tail -30 gpurun_out/w_tests.log | cut -c1-600
