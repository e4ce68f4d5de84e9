# Round 2, call B: GPU tests on the new attention (deferred split-KV merge), bench, reference CUDA baseline, launch list.
set -x
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -x -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/r2b_tests.log 2>&1; tail -6 gpurun_out/r2b_tests.log | cut -c1-300
( timeout 300 python bench.py --no-cpu-baseline ) > gpurun_out/r2b_bench.log 2>&1; tail -1 gpurun_out/r2b_bench.log | cut -c1-900
bash scripts/gpu_ref_cuda_baseline.sh
CMD="python bench.py --steps 2 --warmup 3 --preheat 0 --no-cpu-baseline --no-graph"
timeout 300 $CMD > gpurun_out/plain_r2b.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemv|attn|topk|sampling|fold|embedding' -s 830 -c 340 --csv --log-file gpurun_out/launches_r2b.csv $CMD > gpurun_out/ncu1.log 2>&1
tail -n 3 gpurun_out/ncu1.log
python scripts/launch_summary.py gpurun_out/launches_r2b.csv 12 > gpurun_out/launches_r2b.txt 2>&1; head -12 gpurun_out/launches_r2b.txt
