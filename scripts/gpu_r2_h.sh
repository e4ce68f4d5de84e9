set -x
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -x -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/r2i_tests.log 2>&1; tail -5 gpurun_out/r2i_tests.log | cut -c1-400
run() { tag=$1; shift; timeout 300 python bench.py --no-cpu-baseline --steps 128 --regions 3 "$@" > gpurun_out/r2i_$tag.log 2>&1; echo "$tag: $(grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2i_$tag.log | head -4 | tr '\n' ' ')"; }
run b1
run b2 --batch 2
run b4 --batch 4
run b8 --batch 8
run b16 --batch 16
run fp8_b1 --wformat fp8
run int4_b1 --wformat int4
B200_X_Q=1 run fp8_b1_q --wformat fp8
B200_X_Q=1 run int4_b1_q --wformat int4
run fp8_b16 --wformat fp8 --batch 16
run int4_b16 --wformat int4 --batch 16
run 70b_rank_b8 --config 70b-tp8-rank --batch 8
