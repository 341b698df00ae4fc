#!/bin/bash
# evidence for profiles/: launch list of the bench command, full ncu captures of the stream kernel on C2 / C3 / C5
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-latency > gpurun_out/r02_launches.log 2>&1
for c in C2 C3 C5; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:acb_stream -s 3 -c 1 -f -o gpurun_out/r02_stream_$c python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-latency > gpurun_out/r02_ncu_$c.log 2>&1
done
ls -la gpurun_out/r02_*
