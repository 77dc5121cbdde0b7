#!/bin/bash
# bench.py both arms (as the driver runs them) + launch list + DRAM traffic of one bench launch
mkdir -p gpurun_out
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
( time timeout 600 python bench.py ) > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
( time timeout 600 python bench.py --steps 5 --warmup 3 --no-extras ) > gpurun_out/bench_c_short.json 2> gpurun_out/bench_c_short.err
AGF_NO_WARM=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c.csv python bench.py --steps 3 --warmup 1 --no-extras > gpurun_out/launches_c.log 2>&1
AGF_NO_WARM=1 timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:step_kernel -s 1 -c 1 --csv --log-file gpurun_out/traffic_fp32_uwb_131072_500.csv python profiles/prof_step.py fp32 uwb 131072 500 2 > gpurun_out/traffic.log 2>&1
echo done
