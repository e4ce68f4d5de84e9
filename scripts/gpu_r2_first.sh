# Round 2, first GPU call: the three things round 1 built but never ran on a B200 --
# (1) the reference's own layer sources on the shim launchers, (2) the reference's own CUDA decoder timed next to ours,
# (3) compute-sanitizer (memcheck / racecheck / initcheck / synccheck) over a bounded subset of the parity tests.
#   usage: gpurun --timeout 1500 -- bash scripts/gpu_r2_first.sh
set -x
mkdir -p gpurun_out
B200_RUN_REF_LAYERS=1 timeout -k 5 400 python -m pytest tests/test_reference_programs.py -q -m gpu -k reference_layer_sources -rs -s --timeout 300 -p no:cacheprovider > gpurun_out/r2_ref_layers.log 2>&1; tail -12 gpurun_out/r2_ref_layers.log | cut -c1-400
bash scripts/gpu_ref_cuda_baseline.sh
CS=/usr/local/cuda/bin/compute-sanitizer
PY="python -m pytest -x -q -p no:cacheprovider --timeout 600"
leg() {  # $1 = tag, $2 = tool options, $3... = pytest selection
    tag=$1; opts=$2; shift 2
    timeout -k 10 240 $CS $opts --error-exitcode 0 --print-limit 20 $PY "$@" > gpurun_out/sanitize_$tag.log 2>&1
    echo "rc=$?" >> gpurun_out/sanitize_$tag.log
    { grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" gpurun_out/sanitize_$tag.log | tail -6; grep -m 12 -E "Invalid|Race reported|Uninitialized|hazard" gpurun_out/sanitize_$tag.log; } > gpurun_out/sanitize_$tag.txt
    cat gpurun_out/sanitize_$tag.txt | cut -c1-240
}
leg memcheck_engine   "--tool memcheck"  tests/test_decoder_engine.py -m gpu -k "matches_oracle and (f32 or bf16)"
leg memcheck_ops      "--tool memcheck"  tests/test_ops_gpu.py -m gpu -k "rmsnorm or linear_quantised or decode_mha or topk or sampling or context_attention"
leg racecheck_engine  "--tool racecheck --racecheck-report analysis" tests/test_decoder_engine.py -m gpu -k "test_engine_7b_single_layer_fp32_config0 or test_lm_head_topk_sampling_tail"
leg initcheck_engine  "--tool initcheck" tests/test_decoder_engine.py -m gpu -k "test_engine_matches_oracle and bf16"
leg synccheck_engine  "--tool synccheck" tests/test_decoder_engine.py -m gpu -k "test_engine_7b_single_layer_fp32_config0"
