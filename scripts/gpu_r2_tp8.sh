# Round 2, the 8-GPU session (charged 8x: keep it short).  Tensor-parallel parity at 8 and 4 ranks, then the two scaling tables:
# Llama-2-7B bf16 batch 1 ctx 1024 (BASELINE configs[1] under TP) and Llama-2-70B-shaped bf16 batch 8 ctx 1024 (configs[4]) at TP-2 / 4 / 8.
#   usage: gpurun --gpus 8 --timeout 1500 -- bash scripts/gpu_r2_tp8.sh
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4; do
  timeout 300 $TR --nproc-per-node $n --master-port 2950$n tests/tp_engine_check.py > gpurun_out/r2_tpcheck_$n.log 2>&1; echo "tp_engine_check $n ranks rc=$?"; grep -c ": OK" gpurun_out/r2_tpcheck_$n.log; grep -E "FAILED|Error" gpurun_out/r2_tpcheck_$n.log | head -5
done
bench() { tag=$1; n=$2; shift 2; timeout 400 $TR --nproc-per-node $n --master-port 2960$n bench.py --gpus $n --steps 200 --warmup 8 --regions 5 --no-cpu-baseline "$@" > gpurun_out/r2_$tag.log 2>&1; echo "$tag: rc=$? $(grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms_per_step": [0-9.]*\|"ok": [a-z]*' gpurun_out/r2_$tag.log | head -3 | tr '\n' ' ')"; }
bench 7b_tp8 8
bench 7b_tp4 4
bench 7b_tp2 2
bench 70b_tp8_b8 8 --config 70b --batch 8
bench 70b_tp4_b8 4 --config 70b --batch 8
bench 70b_tp2_b8 2 --config 70b --batch 8
B200_TP_NCCL=1 bench 7b_tp8_nccl 8
