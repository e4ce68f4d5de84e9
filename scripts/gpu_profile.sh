# Launch list + full ncu captures of the two dominant kernels of the decode step (bench.py workload, eager launches).
set -x
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 3 --preheat 0 --no-cpu-baseline --no-graph"
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemv|attn|topk|sampling|fold|embedding' -s 830 -c 340 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_nk_kernel -s 640 -c 4 -o gpurun_out/prof_gemv_$TAG $CMD > gpurun_out/ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_attn_kernel -s 160 -c 1 -o gpurun_out/prof_attn_$TAG $CMD > gpurun_out/ncu3.log 2>&1
tail -n 3 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log
