# chained GEMV kernel: parity tests, then A/B bench on the same box
set -x
timeout -k 5 300 python -m pytest tests/test_decoder_engine.py -x -q -m gpu --timeout 120 -p no:cacheprovider > gpurun_out/chain_tests.log 2>&1; tail -15 gpurun_out/chain_tests.log | cut -c1-300
( timeout 200 python bench.py --no-cpu-baseline ) > gpurun_out/chain_bench.log 2>&1; tail -c 1800 gpurun_out/chain_bench.log
( B200_NO_CHAIN=1 timeout 200 python bench.py --no-cpu-baseline ) > gpurun_out/nochain_bench.log 2>&1; tail -c 1800 gpurun_out/nochain_bench.log
