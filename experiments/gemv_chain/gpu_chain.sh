# chained GEMV kernel: parity tests, then A/B bench on the same box
set -x
export B200_CHAIN=1
timeout -k 5 300 python -m pytest tests/test_decoder_engine.py -x -q -m gpu --timeout 120 -p no:cacheprovider > gpurun_out/chain_tests.log 2>&1; tail -15 gpurun_out/chain_tests.log | cut -c1-300
( timeout 200 python bench.py --no-cpu-baseline ) > gpurun_out/chain_bench.log 2>&1; tail -c 1800 gpurun_out/chain_bench.log
timeout 250 python scripts/chain_trace.py > gpurun_out/chain_trace.log 2>&1; tail -60 gpurun_out/chain_trace.log
