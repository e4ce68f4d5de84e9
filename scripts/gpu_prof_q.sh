set -x
WF=${1:-int4}
CMD="python bench.py --wformat $WF --steps 2 --warmup 3 --preheat 0 --no-cpu-baseline --no-graph"
timeout 300 $CMD > gpurun_out/plain_q.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemv_q_kernel -s 640 -c 4 -o gpurun_out/prof_gemvq_$WF $CMD > gpurun_out/ncu_q.log 2>&1
tail -3 gpurun_out/ncu_q.log
