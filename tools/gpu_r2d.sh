#!/bin/bash
# round 2, pair kernel: parity of the dense / random cases, then kernel-time variants, then the full gpu suite
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_batches or pathological or dense or golden" 2>&1 | tail -15 ) > gpurun_out/r2d_pytest.log
cat gpurun_out/r2d_pytest.log
if grep -q "passed" gpurun_out/r2d_pytest.log && ! grep -q "failed\|error" gpurun_out/r2d_pytest.log; then
  tools/gpu_variants.sh "" noprobe nosurv nodrain 2>&1 | tee gpurun_out/r2d_variants.log
fi
