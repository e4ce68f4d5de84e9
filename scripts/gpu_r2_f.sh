set -x
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -x -q -m gpu -s --timeout 900 -p no:cacheprovider > gpurun_out/r2f_tests.log 2>&1; grep -E "7B 32-layer|7B-shaped|prefill 2048|passed|failed" gpurun_out/r2f_tests.log | cut -c1-400 | head -20
for v in 0 1 0 1; do
  if [ "$v" = "1" ]; then export B200_X_FULL=1; else unset B200_X_FULL; fi
  timeout 200 python bench.py --no-cpu-baseline --steps 256 > gpurun_out/r2f_bench_full$v.log 2>&1
  echo "full $v: $(grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*\|"avg_launch_us": [0-9.]*' gpurun_out/r2f_bench_full$v.log | head -4 | tr '\n' ' ')"
done
unset B200_X_FULL
timeout 200 python bench.py --no-cpu-baseline --steps 128 --batch 4 > gpurun_out/r2f_bench_b4.log 2>&1; echo "b4: $(grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*' gpurun_out/r2f_bench_b4.log | head -2 | tr '\n' ' ')"
B200_X_FULL=1 timeout 200 python bench.py --no-cpu-baseline --steps 128 --batch 4 > gpurun_out/r2f_bench_b4_full.log 2>&1; echo "b4 full: $(grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*' gpurun_out/r2f_bench_b4_full.log | head -2 | tr '\n' ' ')"
