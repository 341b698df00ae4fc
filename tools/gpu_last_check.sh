#!/bin/bash
( timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_iter_long_set.py -m gpu -x -q -k "random_batches or pathological or dense or pair_kernel or iter_long or full_size_c2" 2>&1 | tail -3 )
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-latency 2>&1 | python tools/kline.py "bench C2"
