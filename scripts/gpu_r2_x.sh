# Round 2, session 2: serving throughput (continuous batching over the paged cache against static batching) + paged / batcher / GEMM tests.
set -x
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_batcher.py tests/test_ops_gpu.py -q -m gpu -k "paged or batcher or linear or swiglu" --timeout 300 -p no:cacheprovider > gpurun_out/r2x_tests.log 2>&1; tail -5 gpurun_out/r2x_tests.log | cut -c1-400
timeout 600 python bench.py --mode serve --batch 16 --requests 64 --no-cpu-baseline > gpurun_out/r2x_serve_b16.log 2>&1; tail -c 1500 gpurun_out/r2x_serve_b16.log
timeout 600 python bench.py --mode serve --batch 32 --requests 96 --no-cpu-baseline > gpurun_out/r2x_serve_b32.log 2>&1; tail -c 1500 gpurun_out/r2x_serve_b32.log
