# packed rounding in the norm / GEMV prologues + batcher rejection test: parity + B=1 timing
set -x
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_ops_gpu.py tests/test_decoder_engine.py tests/test_batcher.py tests/test_fullsize_gpu.py -q -m gpu -k "norm or engine or batcher or linear or step" --timeout 600 -p no:cacheprovider > gpurun_out/r2aa_tests.log 2>&1; tail -5 gpurun_out/r2aa_tests.log | cut -c1-400
run() { tag=$1; shift; timeout 300 python bench.py --no-cpu-baseline --steps 256 --regions 5 "$@" > gpurun_out/r2aa_$tag.log 2>&1; echo "$tag: $(grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2aa_$tag.log | head -4 | tr '\n' ' ')"; }
run b1
run b8 --batch 8
timeout 300 python bench.py --mode prefill --prefill-tokens 2048 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2aa_prefill.log 2>&1; grep -o '"value": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2aa_prefill.log | head -2 | tr '\n' ' '
