#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 2500 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --mode weak --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_n${N}_weak.json 2> gpurun_out/bench_n${N}_weak.err
python tools/kline.py "weak N=$N" < gpurun_out/bench_n${N}_weak.json; tail -3 gpurun_out/bench_n${N}_weak.err
