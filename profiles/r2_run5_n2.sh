#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
( timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 profiles/dev_nccl_stages.py ) > $O/nccl_stages_graph2.log 2>&1; echo "rc=$?" >> $O/nccl_stages_graph2.log
( time timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 ) > $O/bench_n2_d.json 2> $O/bench_n2_d.err
( time timeout 240 python -m pytest tests/test_host_facade.py -m gpu -q -s -k monte_carlo ) > $O/gpu_tests_fleet_n2b.log 2>&1; echo "tests rc=$?" >> $O/gpu_tests_fleet_n2b.log
echo done
