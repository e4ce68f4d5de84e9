# Round 2, session 2: ragged batches, TMEM-resident context attention, prefill-shaped norm instances.
set -x
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_ragged.py tests/test_generate.py -x -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/r2r_ragged.log 2>&1; tail -15 gpurun_out/r2r_ragged.log | cut -c1-400
timeout -k 5 600 python -m pytest tests/test_ops_gpu.py tests/test_decoder_engine.py tests/test_fullsize_gpu.py -q -m gpu -k "context_attention or norm or prefill" --timeout 300 -p no:cacheprovider > gpurun_out/r2r_ctx.log 2>&1; tail -25 gpurun_out/r2r_ctx.log | cut -c1-400
timeout 300 python bench.py --mode prefill --prefill-tokens 2048 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2r_prefill.log 2>&1; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2r_prefill.log | head -3 | tr '\n' ' '
CMD="python bench.py --mode prefill --prefill-tokens 2048 --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm|attn|norm|rope|concat|silu|padding|residual|seq_offset' -s 1000 -c 300 --csv --log-file gpurun_out/r2r_launches_prefill.csv $CMD > gpurun_out/r2r_ncu_p.log 2>&1
python scripts/launch_summary.py gpurun_out/r2r_launches_prefill.csv 12 > gpurun_out/r2r_launches_prefill.txt 2>&1; head -24 gpurun_out/r2r_launches_prefill.txt
