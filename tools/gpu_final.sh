#!/bin/bash
# round-end check on one B200: the whole gpu suite, smoke(), the default bench line, the reference arm, then ncu
# captures of the stream kernel on C5 / C3 (profiles/)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q --durations=4 2>&1 | tail -10 ) > gpurun_out/final_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) > gpurun_out/final_smoke.log
timeout 600 python bench.py > gpurun_out/final_bench_C2.json 2> gpurun_out/final_bench_C2.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/final_bench_reference_arm.json 2> gpurun_out/final_ref.err
cat gpurun_out/final_pytest.log gpurun_out/final_smoke.log; python tools/kline.py "bench C2" < gpurun_out/final_bench_C2.json
for c in C5 C3; do
  timeout 240 ncu --set full --import-source on --clock-control none -k regex:acb_stream -s 3 -c 1 -f -o gpurun_out/final_stream_$c python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-latency > gpurun_out/final_ncu_$c.log 2>&1
done
ls -la gpurun_out/final_*
