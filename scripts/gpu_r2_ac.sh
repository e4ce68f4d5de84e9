# context attention: one in four exponentials on the FMA pipe (degree-4 polynomial)
set -x
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_ops_gpu.py tests/test_decoder_engine.py tests/test_fullsize_gpu.py tests/test_batcher.py -q -m gpu -k "context_attention or prefill" --timeout 300 -p no:cacheprovider > gpurun_out/r2ac_tests.log 2>&1; tail -4 gpurun_out/r2ac_tests.log | cut -c1-400
timeout 300 python bench.py --mode prefill --prefill-tokens 2048 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2ac_prefill.log 2>&1; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2ac_prefill.log | head -3 | tr '\n' ' '
CMD="python bench.py --mode prefill --prefill-tokens 2048 --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'attn' -s 100 -c 8 --csv --log-file gpurun_out/r2ac_launches_ctx.csv $CMD > gpurun_out/r2ac_ncu_p.log 2>&1
grep -o 'context_attn_tc_kernel.*' gpurun_out/r2ac_launches_ctx.csv | awk -F'","' '{print $NF}' | head -8 | tr '\n' ' '
