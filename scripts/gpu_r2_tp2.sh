set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_tp_gloo.py -m gpu -q -s --timeout 500 -p no:cacheprovider > gpurun_out/r2_test_tp2.log 2>&1; grep -E "rank|passed|failed|Error" gpurun_out/r2_test_tp2.log | tail -20 | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 2 --steps 256 --warmup 8 --no-cpu-baseline > gpurun_out/r2_bench_tp2.log 2>&1; tail -c 1800 gpurun_out/r2_bench_tp2.log
