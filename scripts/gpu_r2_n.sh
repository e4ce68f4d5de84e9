set -x
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -x -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/r2n_tests.log 2>&1; tail -3 gpurun_out/r2n_tests.log | cut -c1-300
run() { tag=$1; shift; timeout 300 python bench.py --no-cpu-baseline --steps 128 --regions 3 "$@" > gpurun_out/r2n_$tag.log 2>&1; echo "$tag: $(grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2n_$tag.log | head -4 | tr '\n' ' ')"; }
run int4_b8 --wformat int4 --batch 8
run int4_b16 --wformat int4 --batch 16
run fp8_b16 --wformat fp8 --batch 16
