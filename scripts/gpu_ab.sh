set -x
timeout 300 python -m pytest tests/test_decoder_engine.py -m gpu -q -x --timeout 200 -p no:cacheprovider 2>&1 | tail -3
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --steps 128 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('L2 prefetch   ', d['value'], d['ms_per_step'])"
B200_NO_L2_PREFETCH=1 timeout 300 python bench.py --no-cpu-baseline --steps 128 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no L2 prefetch', d['value'], d['ms_per_step'])"
done
