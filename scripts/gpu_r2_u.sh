# Round 2, session 2: paged KV cache + continuous batching (kernels bit-identical to the contiguous ones, iteration loop), context
# attention with the early prologue loads, decode sanity benches after the norm-kernel instance split.
set -x
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_batcher.py tests/test_ragged.py -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/r2u_batcher.log 2>&1; tail -30 gpurun_out/r2u_batcher.log | cut -c1-500
timeout -k 5 600 python -m pytest tests/test_ops_gpu.py tests/test_decoder_engine.py tests/test_generate.py -q -m gpu -k "context_attention or prefill or norm or generate or engine" --timeout 300 -p no:cacheprovider > gpurun_out/r2u_tests.log 2>&1; tail -6 gpurun_out/r2u_tests.log | cut -c1-500
timeout 300 python bench.py --mode prefill --prefill-tokens 2048 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_prefill.log 2>&1; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2u_prefill.log | head -3 | tr '\n' ' '
run() { tag=$1; shift; timeout 300 python bench.py --no-cpu-baseline --steps 128 --regions 3 "$@" > gpurun_out/r2u_$tag.log 2>&1; echo "$tag: $(grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2u_$tag.log | head -4 | tr '\n' ' ')"; }
run b1
run b8 --batch 8
CMD="python bench.py --mode prefill --prefill-tokens 2048 --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'attn' -s 100 -c 6 --csv --log-file gpurun_out/r2u_launches_ctx.csv $CMD > gpurun_out/r2u_ncu_p.log 2>&1
grep -o 'context_attn_tc_kernel.*' gpurun_out/r2u_launches_ctx.csv | awk -F'","' '{print $NF}' | head -6 | tr '\n' ' '
