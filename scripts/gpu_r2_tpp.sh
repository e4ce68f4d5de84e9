# tensor-parallel prefill check on 2 GPUs (tests/tp_engine_check.py: decode legs + b200_decoder_prefill_tp against the un-sharded oracle)
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 2 --master-port 29502 tests/tp_engine_check.py > gpurun_out/r2tpp_tpcheck_2.log 2>&1; echo "tp_engine_check 2 ranks rc=$?"; grep -c ": OK" gpurun_out/r2tpp_tpcheck_2.log; grep -i "prefill\|FAILED\|Error\|error" gpurun_out/r2tpp_tpcheck_2.log | head -12 | cut -c1-400
