set -x
timeout -k 5 300 python -m pytest tests/test_decoder_engine.py tests/test_ops_gpu.py -m gpu -q -k "prefill or context_attention" --timeout 200 --timeout-method=thread -p no:cacheprovider > gpurun_out/test_prefill.log 2>&1; tail -30 gpurun_out/test_prefill.log | cut -c1-300
