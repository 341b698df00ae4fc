#!/bin/bash
# the driver's bench lines (ours + reference arm) and the other configs' kernel lines
mkdir -p gpurun_out
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 python bench.py > gpurun_out/bench_C2.json 2> gpurun_out/bench_C2.err
for c in C3 C4 C5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
done
for f in ref C2 C3 C4 C5; do echo "== $f"; tail -c 1500 gpurun_out/bench_$f.json; tail -3 gpurun_out/bench_$f.err; done
