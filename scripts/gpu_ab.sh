set -x
for i in 1 2; do
B200_L2_PREFETCH=1 timeout 300 python bench.py --no-cpu-baseline --steps 128 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('L2 prefetch   ', d['value'], d['ms_per_step'])"
timeout 300 python bench.py --no-cpu-baseline --steps 128 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no L2 prefetch', d['value'], d['ms_per_step'])"
done
