#!/bin/bash
# round 2, 8-GPU call: the bench as the driver launches it (C3 shard per GPU, 1 M vehicles) and C4 over 8 GPUs (16 M vehicles)
mkdir -p gpurun_out/r2
O=gpurun_out/r2
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv,noheader > $O/smi_n8.txt 2>&1
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 5 ) > $O/bench_n8_l.json 2> $O/bench_n8_l.err
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 8 --steps 20 --warmup 5 --config c4 ) > $O/bench_c4_n8_l.json 2> $O/bench_c4_n8_l.err
( time timeout 200 python bench.py --steps 20 --warmup 5 --no-extras ) > $O/bench_n1_l.json 2> $O/bench_n1_l.err
tail -c 600 $O/bench_n8_l.err; head -c 400 $O/bench_n8_l.json; echo; head -c 400 $O/bench_c4_n8_l.json; echo; head -c 300 $O/bench_n1_l.json
