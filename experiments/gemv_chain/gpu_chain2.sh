set -x
export B200_CHAIN=1
for d in 0 4 8 16; do
( B200_CHAIN_L2_AHEAD=$d timeout 200 python bench.py --no-cpu-baseline --steps 128 ) > gpurun_out/chain_bench_l2_$d.log 2>&1; grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"\|"ms": [0-9.]*' gpurun_out/chain_bench_l2_$d.log
done
B200_CHAIN_L2_AHEAD=8 timeout 250 python scripts/chain_trace.py > gpurun_out/chain_trace.log 2>&1; grep -B1 -A46 "^SMs" gpurun_out/chain_trace.log | grep "phase\|span\|staging"
