# compute-sanitizer over a bounded subset of the parity tests (SURVEY.md section 5 "race detection / sanitizers"): memcheck on the decode engine,
# the prefill engine and the launcher-level ops; racecheck (shared-memory hazards: the mbarrier rings, the reducer hand-off) and initcheck
# (uninitialised global reads: scratch, split-KV partials, exchange buffers) on the smallest decode-engine case.  The sanitizer slows kernels
# 10-100x, so every leg has its own timeout and the summaries (not the full logs) go to gpurun_out/sanitize_*.txt.
#   usage: gpurun --timeout 1500 -- bash scripts/gpu_sanitize.sh
set -x
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
PY="python -m pytest -x -q -p no:cacheprovider --timeout 600"
leg() {  # $1 = tag, $2 = tool options, $3... = pytest selection
    tag=$1; opts=$2; shift 2
    timeout -k 10 420 $CS $opts --error-exitcode 0 --print-limit 20 $PY "$@" > gpurun_out/sanitize_$tag.log 2>&1
    echo "rc=$?" >> gpurun_out/sanitize_$tag.log
    { grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" gpurun_out/sanitize_$tag.log | tail -6; grep -m 12 -E "Invalid|Race reported|Uninitialized|hazard" gpurun_out/sanitize_$tag.log; } > gpurun_out/sanitize_$tag.txt
    cat gpurun_out/sanitize_$tag.txt | cut -c1-240
}
leg memcheck_engine   "--tool memcheck"  tests/test_decoder_engine.py -m gpu -k "matches_oracle and (f32 or bf16)"
leg memcheck_ops      "--tool memcheck"  tests/test_ops_gpu.py -m gpu -k "rmsnorm or linear_quantised or decode_mha or topk or sampling or context_attention"
leg memcheck_generate "--tool memcheck"  tests/test_generate.py -m gpu
leg racecheck_engine  "--tool racecheck --racecheck-report analysis" tests/test_decoder_engine.py -m gpu -k "test_engine_7b_single_layer_fp32_config0 or test_lm_head_topk_sampling_tail"
leg initcheck_engine  "--tool initcheck" tests/test_decoder_engine.py -m gpu -k "test_engine_matches_oracle and bf16"
leg synccheck_engine  "--tool synccheck" tests/test_decoder_engine.py -m gpu -k "test_engine_7b_single_layer_fp32_config0"
