set -x
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --no-cpu-baseline --steps 256 > gpurun_out/r2g_$tag.log 2>&1; echo "$tag: $(grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*' gpurun_out/r2g_$tag.log | head -2 | tr '\n' ' ')"; }
run base A=1
run attn1 B200_X_ATTN1=1
run ohalf B200_X_OHALF=1
run both B200_X_ATTN1=1 B200_X_OHALF=1
run base2 A=1
run both2 B200_X_ATTN1=1 B200_X_OHALF=1
timeout -k 5 600 python -m pytest tests/test_fullsize_gpu.py -x -q -m gpu -s --timeout 900 -p no:cacheprovider -k "batch32 or prefill" > gpurun_out/r2g_tests.log 2>&1; grep -E "prefill 2048|passed|failed" gpurun_out/r2g_tests.log | cut -c1-300
