# quick check: engine + gemv tests, then the default bench
set -x
timeout -k 5 600 python -m pytest tests/test_decoder_engine.py -x -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/quick_tests.log 2>&1; tail -3 gpurun_out/quick_tests.log
( time timeout 300 python bench.py ) > gpurun_out/quick_bench.log 2>&1; tail -c 2500 gpurun_out/quick_bench.log
