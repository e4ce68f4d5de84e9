set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --preheat 0 --no-cpu-baseline --no-graph --regions 1 --batch 8"
timeout 300 $CMD > gpurun_out/plain_r2k.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemv_mma_kernel -s 40 -c 4 -o gpurun_out/prof_mma_b8 $CMD > gpurun_out/ncu_k.log 2>&1
tail -3 gpurun_out/ncu_k.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemv|attn|norm|topk|sampling|fold|embedding' -s 300 -c 460 --csv --log-file gpurun_out/launches_r2k_b8.csv $CMD > gpurun_out/ncu_k2.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_r2k_b8.csv 16 > gpurun_out/launches_r2k_b8.txt 2>&1; head -30 gpurun_out/launches_r2k_b8.txt
