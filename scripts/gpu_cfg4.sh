set -x
for wf in fp8 int4; do for b in 1 16; do
timeout 600 python bench.py --wformat $wf --batch $b --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${wf}_b$b.log 2>&1; tail -c 700 gpurun_out/bench_${wf}_b$b.log
done; done
