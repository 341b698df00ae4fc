#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/published_benchmark.py --words 1000000 --flavour bytes --reference > gpurun_out/published_1m.json 2> gpurun_out/published_1m.err
tail -c 1800 gpurun_out/published_1m.json; tail -3 gpurun_out/published_1m.err
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_on_gpu and unit_search" 2>&1 | tail -2
