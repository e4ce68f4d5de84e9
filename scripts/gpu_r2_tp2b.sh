set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29502 tests/tp_engine_check.py > gpurun_out/r2b_tpcheck_2.log 2>&1; echo "tp_engine_check 2 ranks rc=$?"; grep -c ": OK" gpurun_out/r2b_tpcheck_2.log
bench() { tag=$1; n=$2; shift 2; timeout 400 $TR --nproc-per-node $n --master-port 2960$n bench.py --gpus $n --steps 200 --warmup 8 --regions 5 --no-cpu-baseline "$@" > gpurun_out/r2b_$tag.log 2>&1; echo "$tag: rc=$? $(grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*\|"ok": [a-z]*' gpurun_out/r2b_$tag.log | head -3 | tr '\n' ' ')"; }
bench 70b_tp2_b8 2 --config 70b --batch 8
bench 7b_tp2 2
timeout 200 python bench.py --no-cpu-baseline --steps 128 --regions 3 --batch 8 > gpurun_out/r2b_b8.log 2>&1; echo "b8: $(grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*' gpurun_out/r2b_b8.log | head -2 | tr '\n' ' ')"
