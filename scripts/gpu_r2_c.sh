set -x
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -x -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/r2c_tests.log 2>&1; tail -6 gpurun_out/r2c_tests.log | cut -c1-300
for v in "16 0" "8 0" "4 0" "2 0" "16 1" "8 1" "4 1"; do
  set -- $v
  if [ "$2" = "1" ]; then export B200_X_NODEFER=1; else unset B200_X_NODEFER; fi
  B200_X_MAXSPLIT=$1 timeout 200 python bench.py --no-cpu-baseline --steps 256 > gpurun_out/r2c_bench_$1_$2.log 2>&1
  echo "maxsplit $1 nodefer $2: $(grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*' gpurun_out/r2c_bench_$1_$2.log | head -2 | tr '\n' ' ')"
done
