# Round 2, session 2: split-KV merge through flagged words (no fence / ticket inside the engine): parity + B=1 / B=8 / GQA timing.
set -x
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_ops_gpu.py tests/test_decoder_engine.py tests/test_ragged.py tests/test_batcher.py tests/test_generate.py tests/test_fullsize_gpu.py -q -m gpu -k "mha or engine or ragged or paged or batcher or generate or step" --timeout 600 -p no:cacheprovider > gpurun_out/r2y_tests.log 2>&1; tail -8 gpurun_out/r2y_tests.log | cut -c1-400
run() { tag=$1; shift; timeout 300 python bench.py --no-cpu-baseline --steps 256 --regions 5 "$@" > gpurun_out/r2y_$tag.log 2>&1; echo "$tag: $(grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2y_$tag.log | head -4 | tr '\n' ' ')"; }
run b1
run b8 --batch 8
run 70b_rank_b8 --config 70b-tp8-rank --batch 8 --steps 64
run b32 --batch 32 --ctx 2048 --steps 32 --regions 3
CMD="python bench.py --steps 2 --warmup 3 --preheat 0 --no-cpu-baseline --no-graph --regions 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemv|attn|topk|sampling|fold|embedding|norm' -s 830 -c 340 --csv --log-file gpurun_out/r2y_launches_b1.csv $CMD > gpurun_out/r2y_ncu1.log 2>&1
python scripts/launch_summary.py gpurun_out/r2y_launches_b1.csv 12 > gpurun_out/r2y_launches_b1.txt 2>&1; head -14 gpurun_out/r2y_launches_b1.txt
