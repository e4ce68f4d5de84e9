# Round 2, second 8-GPU session: after the wide norm kernel (LL reduce with every load of a row in flight) and the 16-byte LL pushes of the
# tensor-core GEMV.  Parity first; the benches only run if it is green.
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29508 tests/tp_engine_check.py > gpurun_out/r2c_tpcheck_8.log 2>&1; rc=$?; echo "tp_engine_check 8 ranks rc=$rc"; grep -c ": OK" gpurun_out/r2c_tpcheck_8.log
if [ $rc -ne 0 ]; then grep -E "FAILED|Error" gpurun_out/r2c_tpcheck_8.log | head; exit 1; fi
bench() { tag=$1; n=$2; shift 2; timeout 300 $TR --nproc-per-node $n --master-port 2960$n bench.py --gpus $n --steps 200 --warmup 8 --regions 5 --no-cpu-baseline "$@" > gpurun_out/r2c_$tag.log 2>&1; echo "$tag: rc=$? $(grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*\|"ok": [a-z]*' gpurun_out/r2c_$tag.log | head -3 | tr '\n' ' ')"; }
bench 70b_tp8_b8 8 --config 70b --batch 8
bench 70b_tp4_b8 4 --config 70b --batch 8
bench 7b_tp8 8
bench 7b_tp4 4
