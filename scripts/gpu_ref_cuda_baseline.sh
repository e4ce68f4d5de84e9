# The reference's OWN CUDA decode path on this GPU next to ours (SURVEY.md 8d "GPU reference baseline"): oracle/_ref/ref_gpu_bench times
# LlamaSelfDecoder<float>::forward (fp32, batch 1, step <= 128: the domain where its kernels are valid), then bench.py times this repo's
# engine on the same shape and step in fp32 and in bf16.  Usage: gpurun --timeout 600 -- bash scripts/gpu_ref_cuda_baseline.sh
set -x
mkdir -p gpurun_out
( timeout -k 5 240 oracle/_ref/ref_gpu_bench 32 128 10 2 > gpurun_out/ref_cuda_7b_step128.json 2> gpurun_out/ref_cuda_7b_step128.err; echo "rc=$?" ) 2>&1 | tail -1
tail -c 1200 gpurun_out/ref_cuda_7b_step128.json
tail -5 gpurun_out/ref_cuda_7b_step128.err | cut -c1-300
# one layer alone (configs[0] shape), more repetitions
( timeout -k 5 120 oracle/_ref/ref_gpu_bench 1 64 50 5 > gpurun_out/ref_cuda_7b_1layer.json 2>/dev/null; echo "rc=$?" ) 2>&1 | tail -1
tail -c 1200 gpurun_out/ref_cuda_7b_1layer.json
# ours at the same context (127 cached positions + the new one), bf16 headline format
timeout -k 5 300 python bench.py --ctx 127 --steps 128 --warmup 8 > gpurun_out/bench_ctx127.log 2>&1; tail -1 gpurun_out/bench_ctx127.log | cut -c1-600
