set -x
timeout -k 5 600 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_r1.log 2>&1; tail -5 gpurun_out/test_r1.log
timeout -k 5 600 python bench.py > gpurun_out/bench_r1.log 2>&1; tail -c 3000 gpurun_out/bench_r1.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemv|attn|topk|sampling|fold|embedding' -s 830 -c 340 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_nk_kernel -s 640 -c 4 -o gpurun_out/prof_gemv_r1 $CMD > gpurun_out/ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_attn_kernel -s 160 -c 1 -o gpurun_out/prof_attn_r1 $CMD > gpurun_out/ncu3.log 2>&1
tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log
