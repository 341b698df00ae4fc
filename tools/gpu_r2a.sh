#!/bin/bash
# round-2 first GPU pass: parity, then kernel-time variants, then one ncu capture
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2a_pytest.log
{
  tools/gpu_variants.sh ":" ":ACB_FILTER=4,1,20,0" "noprobe:"
} > gpurun_out/r2a_variants.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:acb_stream -s 3 -c 1 -o gpurun_out/r2a_stream python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2a_ncu.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python __graft_entry__.py --smoke > gpurun_out/r2a_sanit.log 2>&1
tail -3 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_variants.log; tail -2 gpurun_out/r2a_sanit.log
