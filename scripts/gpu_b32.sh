set -x
( timeout 400 python bench.py --batch 32 --ctx 2048 --steps 32 --no-cpu-baseline ) > gpurun_out/bench_b32.log 2>&1; grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*' gpurun_out/bench_b32.log | head -2
