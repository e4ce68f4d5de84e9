# ncu --set full of the decode attention kernel (flagged-word merge) inside the B=1 step
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --preheat 0 --no-cpu-baseline --no-graph --regions 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_attn_kernel -s 160 -c 1 -o gpurun_out/prof_attn_r2z $CMD > gpurun_out/r2z_ncu.log 2>&1; tail -2 gpurun_out/r2z_ncu.log
