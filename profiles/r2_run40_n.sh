#!/bin/bash
# the bench as the driver launches it on N GPUs of one box: bash profiles/r2_run17_n.sh <N>
N=$1
mkdir -p gpurun_out/r2
O=gpurun_out/r2
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N bench.py --gpus $N --steps 20 --warmup 5 ) > $O/bench_n${N}_final2.json 2> $O/bench_n${N}_final2.err
head -c 300 $O/bench_n${N}_final2.json
