#!/bin/bash
# parity (dense cases first, watchdog build, short timeouts) + kernel-time variants (+ optional ncu capture: NCU=1)
mkdir -p gpurun_out
( ACB_LIB=$PWD/pyahocorasick_b200/_native/libacb200_wd.so timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_batches or pathological or dense" 2>&1 | tail -15 ) > gpurun_out/r2b_pytest.log
if grep -q "passed" gpurun_out/r2b_pytest.log && ! grep -q "failed" gpurun_out/r2b_pytest.log; then
  ( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) >> gpurun_out/r2b_pytest.log
fi
tools/gpu_variants.sh "$@" > gpurun_out/r2b_variants.log 2>&1
if [ -n "$NCU" ]; then
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:acb_stream -s 3 -c 1 -f -o gpurun_out/r2b_stream python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-latency > gpurun_out/r2b_ncu.log 2>&1
fi
tail -6 gpurun_out/r2b_pytest.log; cat gpurun_out/r2b_variants.log
