set -x
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_ops_gpu.py tests/test_decoder_engine.py tests/test_fullsize_gpu.py tests/test_weights.py -x -q -m gpu --timeout 900 -p no:cacheprovider -k "linear or gemv or engine or quantised or weights" > gpurun_out/r2m_tests.log 2>&1; tail -3 gpurun_out/r2m_tests.log | cut -c1-300
run() { tag=$1; shift; timeout 300 python bench.py --no-cpu-baseline --steps 128 --regions 3 "$@" > gpurun_out/r2m_$tag.log 2>&1; echo "$tag: $(grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2m_$tag.log | head -4 | tr '\n' ' ')"; }
run fp8_b2 --wformat fp8 --batch 2
run fp8_b8 --wformat fp8 --batch 8
run fp8_b16 --wformat fp8 --batch 16
run int4_b2 --wformat int4 --batch 2
run int4_b8 --wformat int4 --batch 8
run int4_b16 --wformat int4 --batch 16
B200_X_Q1MMA=1 run fp8_b1_mma --wformat fp8
B200_X_Q1MMA=1 run int4_b1_mma --wformat int4
run b8 --batch 8
run b16 --batch 16
