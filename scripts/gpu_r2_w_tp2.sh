# Round 2, session 2: tensor-parallel sanity on 2 GPUs after the norm-kernel instance split (the TP path has its own instances now).
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29502 tests/tp_engine_check.py > gpurun_out/r2w_tpcheck_2.log 2>&1; echo "tp_engine_check 2 ranks rc=$?"; grep -c ": OK" gpurun_out/r2w_tpcheck_2.log; tail -3 gpurun_out/r2w_tpcheck_2.log | cut -c1-300
bench() { tag=$1; n=$2; shift 2; timeout 400 $TR --nproc-per-node $n --master-port 2960$n bench.py --gpus $n --steps 200 --warmup 8 --regions 3 --no-cpu-baseline "$@" > gpurun_out/r2w_$tag.log 2>&1; echo "$tag: rc=$? $(grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*\|"ok": [a-z]*' gpurun_out/r2w_$tag.log | head -3 | tr '\n' ' ')"; }
bench 7b_tp2 2
bench 70b_tp2_b8 2 --config 70b --batch 8
