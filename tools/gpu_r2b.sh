#!/bin/bash
# parity + kernel-time variants (+ optional ncu capture of the default build: NCU=1)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2b_pytest.log
tools/gpu_variants.sh "$@" > gpurun_out/r2b_variants.log 2>&1
if [ -n "$NCU" ]; then
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:acb_stream -s 3 -c 1 -f -o gpurun_out/r2b_stream python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2b_ncu.log 2>&1
fi
tail -3 gpurun_out/r2b_pytest.log; cat gpurun_out/r2b_variants.log
