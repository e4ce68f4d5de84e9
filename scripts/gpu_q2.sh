set -x
timeout -k 5 400 python -m pytest tests/test_ops_gpu.py tests/test_decoder_engine.py -m gpu -q -k "quantis" --timeout 200 -p no:cacheprovider > gpurun_out/test_q.log 2>&1; tail -4 gpurun_out/test_q.log | cut -c1-300
for wf in int4 fp8; do for b in 1 16; do
  timeout 300 python bench.py --wformat $wf --batch $b --steps 64 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${wf}_b$b.log 2>&1
  python -c "import sys,json; d=json.loads(open('gpurun_out/bench_${wf}_b$b.log').read().strip().splitlines()[-1]); print('$wf b$b', round(d['value'],1),'tok/s', round(d['ms_per_step'],3),'ms', 'linears GB/s', round(d['roofline']['achieved']), 'whole frac', round(d['roofline']['whole_step']['frac'],3))"
done; done
