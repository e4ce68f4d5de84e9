set -x
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 300 python bench.py --no-cpu-baseline --steps 128 --regions 3 "$@" > gpurun_out/r2p_$tag.log 2>&1; echo "$tag: $(grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2p_$tag.log | head -4 | tr '\n' ' ')"; }
for pen in 49152 8192 0; do
  export B200_X_PARTPEN=$pen
  run b8_pen$pen --batch 8
  run b16_pen$pen --batch 16
  run fp8_b16_pen$pen --batch 16 --wformat fp8
done
