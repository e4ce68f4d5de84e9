set -x
nvidia-smi -L
timeout 900 python -m pytest tests/test_tp_gloo.py -m gpu -q -s --timeout 800 -p no:cacheprovider > gpurun_out/test_tp2.log 2>&1; tail -15 gpurun_out/test_tp2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 2 --steps 32 --warmup 4 > gpurun_out/bench_tp2.log 2>&1; tail -c 2500 gpurun_out/bench_tp2.log
