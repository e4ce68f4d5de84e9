set -x
nvidia-smi -L
timeout 600 python -m pytest tests/test_tp_gloo.py -m gpu -q -s --timeout 500 -p no:cacheprovider > gpurun_out/test_tp2.log 2>&1; grep -E "rank|passed|failed" gpurun_out/test_tp2.log | tail -20
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 2 --steps 64 --warmup 4 > gpurun_out/bench_tp2.log 2>&1; tail -c 1500 gpurun_out/bench_tp2.log
B200_TP_NCCL=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --steps 64 --warmup 4 > gpurun_out/bench_tp2_nccl.log 2>&1; tail -c 600 gpurun_out/bench_tp2_nccl.log
