#!/bin/bash
# evidence for profiles/: launch list of the bench command and one full capture of the pair kernel on C2 planted
mkdir -p gpurun_out
timeout 400 ncu --set full --import-source on --clock-control none -k regex:acb_pair -s 3 -c 1 -f -o gpurun_out/prof_pair_C2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-latency > gpurun_out/prof_ncu_C2.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-latency > gpurun_out/prof_launches.log 2>&1
ls -la gpurun_out/prof_*
