set -x
CMD="python bench.py --config 70b-tp8-rank --batch 8 --steps 2 --warmup 3 --preheat 0 --no-cpu-baseline --no-graph"
timeout 300 python bench.py --config 70b-tp8-rank --batch 8 --steps 64 --no-cpu-baseline > gpurun_out/bench_70b_rank.log 2>&1; grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*' gpurun_out/bench_70b_rank.log | head -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 400 --csv --log-file gpurun_out/launches_70b_rank.csv $CMD > gpurun_out/ncu_70b.log 2>&1
tail -2 gpurun_out/ncu_70b.log
