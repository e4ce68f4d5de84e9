set -x
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_decoder_engine.py tests/test_weights.py tests/test_generate.py -x -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/r2q_tests.log 2>&1; tail -4 gpurun_out/r2q_tests.log | cut -c1-300
timeout 300 python bench.py --mode prefill --prefill-tokens 2048 --steps 4 --warmup 3 --no-cpu-baseline --wformat fp8 > gpurun_out/r2q_prefill_fp8.log 2>&1; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/r2q_prefill_fp8.log | head -2 | tr '\n' ' '
timeout 300 python bench.py --mode prefill --prefill-tokens 2048 --steps 4 --warmup 3 --no-cpu-baseline --wformat int4 > gpurun_out/r2q_prefill_int4.log 2>&1; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/r2q_prefill_int4.log | head -2 | tr '\n' ' '
