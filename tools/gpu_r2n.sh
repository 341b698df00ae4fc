#!/bin/bash
# round 2 final: the whole gpu suite, the default bench line, the reference arm, the other single-GPU configs
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -14 ) > gpurun_out/r2n_pytest.log
timeout 600 python bench.py > gpurun_out/r2n_bench_C2.json 2> gpurun_out/r2n_bench_C2.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r2n_bench_reference_arm.json 2> gpurun_out/r2n_ref.err
for c in C3 C4 C5; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/r2n_bench_$c.json 2> gpurun_out/r2n_bench_$c.err; done
timeout 300 python bench.py --variant sparse --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-latency > gpurun_out/r2n_bench_C2_sparse.json 2>/dev/null
cat gpurun_out/r2n_pytest.log
for f in C2 C3 C4 C5 C2_sparse; do python tools/kline.py "bench $f" < gpurun_out/r2n_bench_$f.json; done
tail -c 600 gpurun_out/r2n_bench_reference_arm.json; tail -3 gpurun_out/r2n_bench_C2.err
