set -x
( timeout 200 python bench.py --no-cpu-baseline ) > gpurun_out/b1_bench.log 2>&1; grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*\|"ms": [0-9.]*' gpurun_out/b1_bench.log | head -4
( timeout 200 python bench.py --no-cpu-baseline --wformat fp8 ) > gpurun_out/b1_bench_fp8.log 2>&1; grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*' gpurun_out/b1_bench_fp8.log | head -2
timeout -k 5 300 python -m pytest tests/test_decoder_engine.py -x -q -m gpu --timeout 200 -p no:cacheprovider -k "engine_matches or quantised" > gpurun_out/b1_tests.log 2>&1; tail -2 gpurun_out/b1_tests.log
