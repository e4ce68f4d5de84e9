# Round 2, session 2: stream-K tail in the prefill GEMMs (whole waves of whole tiles, last partial wave shared by all SMs).
set -x
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_ops_gpu.py tests/test_decoder_engine.py tests/test_fullsize_gpu.py tests/test_weights.py tests/test_batcher.py -q -m gpu -k "linear or swiglu or prefill or paged or batcher" --timeout 300 -p no:cacheprovider > gpurun_out/r2v_tests.log 2>&1; tail -12 gpurun_out/r2v_tests.log | cut -c1-600
timeout 300 python bench.py --mode prefill --prefill-tokens 2048 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2v_prefill.log 2>&1; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2v_prefill.log | head -3 | tr '\n' ' '
CMD="python bench.py --mode prefill --prefill-tokens 2048 --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm|attn|norm|rope|concat|silu|padding|residual|seq_offset' -s 1000 -c 260 --csv --log-file gpurun_out/r2v_launches_prefill.csv $CMD > gpurun_out/r2v_ncu_p.log 2>&1
python scripts/launch_summary.py gpurun_out/r2v_launches_prefill.csv 10 > gpurun_out/r2v_launches_prefill.txt 2>&1; head -20 gpurun_out/r2v_launches_prefill.txt
