set -x
timeout -k 5 400 python -m pytest tests/test_decoder_engine.py tests/test_ops_gpu.py -x -q -m gpu --timeout 200 -p no:cacheprovider -k "engine or mha or attn or decode" > gpurun_out/attn_tests.log 2>&1; tail -4 gpurun_out/attn_tests.log | cut -c1-300
( timeout 200 python bench.py --no-cpu-baseline ) > gpurun_out/attn_bench.log 2>&1; grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*' gpurun_out/attn_bench.log | head -2
timeout 300 python bench.py --config 70b-tp8-rank --batch 8 --steps 64 --no-cpu-baseline > gpurun_out/bench_70b_rank.log 2>&1; grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*' gpurun_out/bench_70b_rank.log | head -2
timeout 300 python bench.py --batch 8 --steps 64 --no-cpu-baseline > gpurun_out/bench_b8.log 2>&1; grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*' gpurun_out/bench_b8.log | head -2
