# last call of the round: the driver's GPU sequence on the final tree (full GPU suite, smoke, headline bench)
set -x
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -x -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/r2bl_tests.log 2>&1; tail -3 gpurun_out/r2bl_tests.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2bl_smoke.log 2>&1; tail -2 gpurun_out/r2bl_smoke.log
( time timeout 600 python bench.py ) > gpurun_out/r2bl_bench.log 2>&1; grep '^{"metric"' gpurun_out/r2bl_bench.log | cut -c1-300
