set -x
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 300 python bench.py --no-cpu-baseline --steps 128 --regions 3 "$@" > gpurun_out/r2o_$tag.log 2>&1; echo "$tag: $(grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2o_$tag.log | head -4 | tr '\n' ' ')"; }
export B200_X_Q8=1
run int4_b8_q --wformat int4 --batch 8
run int4_b2_q --wformat int4 --batch 2
run int4_b4_q --wformat int4 --batch 4
run fp8_b8_q --wformat fp8 --batch 8
run fp8_b2_q --wformat fp8 --batch 2
unset B200_X_Q8
run int4_b4 --wformat int4 --batch 4
run fp8_b4 --wformat fp8 --batch 4
