# what the driver does at round end, on one box: GPU tests, smoke, both bench arms
set -x
timeout -k 5 900 python -m pytest tests -x -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/rehearsal_tests.log 2>&1; tail -4 gpurun_out/rehearsal_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/rehearsal_smoke.log 2>&1; tail -3 gpurun_out/rehearsal_smoke.log
( time timeout 400 python bench.py --impl reference --gpus 1 --steps 16 --warmup 3 ) > gpurun_out/rehearsal_ref.log 2>&1; tail -c 900 gpurun_out/rehearsal_ref.log
( time timeout 600 python bench.py ) > gpurun_out/rehearsal_bench.log 2>&1; tail -c 2600 gpurun_out/rehearsal_bench.log
