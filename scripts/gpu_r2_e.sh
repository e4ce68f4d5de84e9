set -x
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -x -q -m gpu -s --timeout 900 -p no:cacheprovider > gpurun_out/r2e_tests.log 2>&1; grep -E "7B 32-layer|prefill 2048|quantisation|passed|failed" gpurun_out/r2e_tests.log | cut -c1-500 | head -20
timeout 200 python bench.py --no-cpu-baseline --steps 256 > gpurun_out/r2e_bench.log 2>&1
echo "bench: $(grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*' gpurun_out/r2e_bench.log | head -3 | tr '\n' ' ')"
