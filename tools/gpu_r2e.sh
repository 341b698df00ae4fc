#!/bin/bash
# ncu captures of the pair kernel: level-1-only build (sparse text) and the full kernel (planted)
mkdir -p gpurun_out
ACB_LIB=$PWD/pyahocorasick_b200/_native/libacb200_nosurv.so timeout 300 ncu --set full --import-source on --clock-control none -k regex:acb_pair -s 3 -c 1 -f -o gpurun_out/r2e_nosurv python bench.py --variant sparse --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-latency > gpurun_out/r2e_nosurv.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:acb_pair -s 3 -c 1 -f -o gpurun_out/r2e_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-latency > gpurun_out/r2e_full.log 2>&1
ls -la gpurun_out/r2e_*
