set -x
timeout -k 5 300 python -m pytest tests/test_generate.py -x -q -m gpu --timeout 120 -p no:cacheprovider > gpurun_out/gen_tests.log 2>&1; tail -30 gpurun_out/gen_tests.log | cut -c1-400
