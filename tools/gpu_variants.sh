#!/bin/bash
# Time compile-time variants of the kernels in ONE gpu call (a call costs ~3 GPU-minutes before the first command runs).
# Build them first, e.g.
#   python -c "from pyahocorasick_b200 import build as B; B.build(force=True, defines=['ACB_PAIR_CONSUMERS=26'], out=B.OUT_DIR+'/libacb200_p26.so')"
# usage: tools/gpu_variants.sh [name:variant:config ...]     name '' = the default library; variant planted|sparse; config C2..C5
mkdir -p gpurun_out
run() { name=$1; var=$2; cfg=$3; lib=$PWD/pyahocorasick_b200/_native/libacb200${name:+_$name}.so
  ACB_LIB=$lib timeout 200 python bench.py --config $cfg --steps 20 --warmup 5 --variant $var --no-cpu-baseline --no-e2e --no-latency 2>&1 | python tools/kline.py "lib=${name:-default} $cfg variant=$var"; }
for spec in "$@"; do
  IFS=: read -r name var cfg <<< "$spec"
  if [ -n "$name" ]; then ( ACB_LIB=$PWD/pyahocorasick_b200/_native/libacb200_$name.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_batches or pathological or dense or pair_kernel" 2>&1 | tail -1 ); fi
  run "$name" "${var:-planted}" "${cfg:-C2}"
done 2>&1 | tee gpurun_out/variants.log
