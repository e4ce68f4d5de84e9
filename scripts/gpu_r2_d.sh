set -x
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_fullsize_gpu.py tests/test_decoder_engine.py tests/test_generate.py -x -q -m gpu -s --timeout 600 -p no:cacheprovider > gpurun_out/r2d_tests.log 2>&1; grep -E "7B 32-layer|prefill 2048|passed|failed|Error|error" gpurun_out/r2d_tests.log | cut -c1-400 | head -20
for v in 0 1 0 1; do
  if [ "$v" = "1" ]; then export B200_X_NOPF=1; else unset B200_X_NOPF; fi
  timeout 200 python bench.py --no-cpu-baseline --steps 256 > gpurun_out/r2d_bench_nopf$v.log 2>&1
  echo "nopf $v: $(grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*\|"avg_launch_us": [0-9.]*' gpurun_out/r2d_bench_nopf$v.log | head -3 | tr '\n' ' ')"
done
