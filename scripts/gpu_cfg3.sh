set -x
timeout 600 python bench.py --mode prefill --prefill-tokens 2048 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_prefill.log 2>&1; tail -c 1500 gpurun_out/bench_prefill.log
B200_CTX_ATTN_SIMT=1 timeout 600 python bench.py --mode prefill --prefill-tokens 2048 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_prefill_simt.log 2>&1; tail -c 600 gpurun_out/bench_prefill_simt.log
timeout 600 python bench.py --batch 32 --ctx 2048 --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b32.log 2>&1; tail -c 1500 gpurun_out/bench_b32.log
timeout 600 python bench.py --batch 8 --ctx 1024 --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b8.log 2>&1; tail -c 800 gpurun_out/bench_b8.log
CMD="python bench.py --mode prefill --prefill-tokens 2048 --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain_prefill.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm|attn|norm|rope|concat|silu|padding|residual|seq_offset' -s 1000 -c 400 --csv --log-file gpurun_out/launches_prefill.csv $CMD > gpurun_out/ncu_p.log 2>&1
tail -n 3 gpurun_out/ncu_p.log
