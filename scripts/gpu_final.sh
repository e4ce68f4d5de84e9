set -x
timeout -k 5 150 python -m pytest tests -x -q -m gpu --timeout 120 -p no:cacheprovider > gpurun_out/final_tests.log 2>&1; tail -4 gpurun_out/final_tests.log | cut -c1-300
