#!/bin/bash
# round 2, 2-GPU call: the NCCL read-out inside the C ABI -- C++ fleet example sharded over 2 GPUs vs unsharded, and the
# bench as the driver launches it at N = 2 (torchrun, one rank per GPU, the library's own communicator).
mkdir -p gpurun_out/r2
O=gpurun_out/r2
nvidia-smi -L > $O/smi_n2.txt 2>&1
( time timeout 600 python -m pytest tests/test_host_facade.py -m gpu -q -s -k monte_carlo ) > $O/gpu_tests_fleet_n2.log 2>&1; echo "tests rc=$?" >> $O/gpu_tests_fleet_n2.log
( time NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 ) > $O/bench_n2_c.json 2> $O/bench_n2_c.err
( time timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras ) > $O/bench_n1_c.json 2> $O/bench_n1_c.err
echo done
