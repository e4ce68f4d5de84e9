# Round 2, session 2, final evidence on one B200: what the driver runs (GPU tests, smoke, both bench arms), the bench lines of the BASELINE
# configurations, the B=1 launch list.  usage: gpurun --timeout 2400 -- bash scripts/gpu_r2b_final.sh
set -x
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -x -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/r2bf_tests.log 2>&1; tail -3 gpurun_out/r2bf_tests.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2bf_smoke.log 2>&1; tail -2 gpurun_out/r2bf_smoke.log
( time timeout 300 python bench.py --impl reference --gpus 1 --steps 4 --warmup 3 ) > gpurun_out/r2bf_ref.log 2>&1; tail -c 400 gpurun_out/r2bf_ref.log | cut -c1-400
( time timeout 600 python bench.py ) > gpurun_out/r2bf_bench.log 2>&1; grep '^{"metric"' gpurun_out/r2bf_bench.log | cut -c1-400
run() { tag=$1; shift; timeout 300 python bench.py --no-cpu-baseline --steps 128 --regions 3 "$@" > gpurun_out/r2bf_$tag.log 2>&1; echo "$tag: $(grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2bf_$tag.log | head -4 | tr '\n' ' ')"; }
run b8 --batch 8
run b16 --batch 16
run fp8_b1 --wformat fp8
run fp8_b16 --wformat fp8 --batch 16
run int4_b1 --wformat int4
run int4_b16 --wformat int4 --batch 16
run b32_ctx2048 --batch 32 --ctx 2048 --steps 32
run 70b_rank_b8 --config 70b-tp8-rank --batch 8
timeout 300 python bench.py --mode prefill --prefill-tokens 2048 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2bf_prefill.log 2>&1; grep -o '"value": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2bf_prefill.log | head -2 | tr '\n' ' '
CMD="python bench.py --steps 2 --warmup 3 --preheat 0 --no-cpu-baseline --no-graph --regions 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemv|attn|topk|sampling|fold|embedding|norm' -s 830 -c 340 --csv --log-file gpurun_out/r2bf_launches_b1.csv $CMD > gpurun_out/r2bf_ncu1.log 2>&1
python scripts/launch_summary.py gpurun_out/r2bf_launches_b1.csv 12 > gpurun_out/r2bf_launches_b1.txt 2>&1; head -8 gpurun_out/r2bf_launches_b1.txt
