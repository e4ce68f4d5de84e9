set -x
CMD="python scripts/bench_linear.py 1 lm_head"
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemv_nk_kernel -s 8 -c 1 -o gpurun_out/prof_gemv3 $CMD > gpurun_out/ncu_gemv3.log 2>&1
tail -3 gpurun_out/ncu_gemv3.log
