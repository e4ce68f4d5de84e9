set -x
timeout 250 python scripts/chain_trace.py > gpurun_out/chain_trace.log 2>&1; tail -60 gpurun_out/chain_trace.log
