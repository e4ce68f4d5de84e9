set -x
timeout -k 5 600 python -m pytest tests/test_reference_programs.py -x -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/gen_tests.log 2>&1; tail -30 gpurun_out/gen_tests.log | cut -c1-500
