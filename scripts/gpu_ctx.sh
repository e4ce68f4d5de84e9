set -x
timeout -k 5 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "context_attention or prefill_chain" --timeout 120 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_ctx.log 2>&1; tail -40 gpurun_out/test_ctx.log | cut -c1-400
