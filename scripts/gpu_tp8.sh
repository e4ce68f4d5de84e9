set -x
nvidia-smi -L | wc -l
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29510 tests/tp_engine_check.py > gpurun_out/test_tp8.log 2>&1; grep -E "rank 0|FAILED|Error" gpurun_out/test_tp8.log | tail -12
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 64 --warmup 4 > gpurun_out/bench_tp8.log 2>&1; tail -c 1200 gpurun_out/bench_tp8.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --config 70b --batch 8 --steps 32 --warmup 4 > gpurun_out/bench_70b_tp8.log 2>&1; tail -c 1500 gpurun_out/bench_70b_tp8.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 64 --warmup 4 > gpurun_out/bench_tp4.log 2>&1; tail -c 700 gpurun_out/bench_tp4.log
