# what the driver does at round end, on one box: GPU tests, smoke, both bench arms
set -x
timeout -k 5 900 python -m pytest tests -x -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/rehearsal_tests.log 2>&1; tail -4 gpurun_out/rehearsal_tests.log
# the reference's own layer sources on the shim's launchers (opt-in until it has passed once on a B200; see tests/test_reference_programs.py)
B200_RUN_REF_LAYERS=1 timeout -k 5 600 python -m pytest tests/test_reference_programs.py -q -m gpu -k reference_layer_sources -rs -s --timeout 300 -p no:cacheprovider > gpurun_out/rehearsal_ref_layers.log 2>&1; tail -12 gpurun_out/rehearsal_ref_layers.log | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/rehearsal_smoke.log 2>&1; tail -3 gpurun_out/rehearsal_smoke.log
( time timeout 400 python bench.py --impl reference --gpus 1 --steps 16 --warmup 3 ) > gpurun_out/rehearsal_ref.log 2>&1; tail -c 900 gpurun_out/rehearsal_ref.log
( time timeout 600 python bench.py ) > gpurun_out/rehearsal_bench.log 2>&1; tail -c 2600 gpurun_out/rehearsal_bench.log
