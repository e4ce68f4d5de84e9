# serve_example (C++ through the C ABI) + the shim programs
set -x
mkdir -p gpurun_out
timeout 120 llm-inference-engine_b200/shim/_own_programs/serve_example > gpurun_out/r2ab_serve_example.log 2>&1; echo "serve_example rc=$?"; cat gpurun_out/r2ab_serve_example.log | tail -8
timeout -k 5 600 python -m pytest tests/test_reference_programs.py -q -m gpu -k "serve or llama_model or chat" --timeout 300 -p no:cacheprovider > gpurun_out/r2ab_tests.log 2>&1; tail -3 gpurun_out/r2ab_tests.log | cut -c1-300
