set -x
timeout 600 python -m pytest tests/test_reference_programs.py -m gpu -q -s --timeout 300 -p no:cacheprovider 2>&1 | grep -E "^(test_|[a-z_]+example|[a-z_]+): shim|passed|failed" > gpurun_out/ref_programs_verdicts.log; cat gpurun_out/ref_programs_verdicts.log
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --steps 128 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cluster   ', d['value'], d['ms_per_step'])"
B200_ATTN_NO_CLUSTER=1 timeout 300 python bench.py --no-cpu-baseline --steps 128 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no-cluster', d['value'], d['ms_per_step'])"
done
