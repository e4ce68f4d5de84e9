set -x
timeout -k 5 300 python -m pytest tests/test_reference_programs.py -x -q -m gpu --timeout 200 -p no:cacheprovider -k "llama_model" > gpurun_out/gen_tests.log 2>&1; tail -30 gpurun_out/gen_tests.log | cut -c1-600
llm-inference-engine_b200/shim/_own_programs/llama_model_example f16 8 | tail -12
