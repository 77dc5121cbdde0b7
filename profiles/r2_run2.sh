#!/bin/bash
# round 2, GPU call 2: population-level parity of the fast kernels after the compensated FP32 integration, every bench arm
# (after the event-ring fix), DRAM traffic of one bench-sized launch, launch list of the bench command.
mkdir -p gpurun_out/r2
O=gpurun_out/r2
( time timeout 1200 python -m pytest tests/test_fast_population_gpu.py -m gpu -q -s ) > $O/gpu_tests_fastpop_b.log 2>&1; echo "tests rc=$?" >> $O/gpu_tests_fastpop_b.log
( time timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_fast_population_gpu.py ) > $O/gpu_tests_b.log 2>&1; echo "tests rc=$?" >> $O/gpu_tests_b.log
( time timeout 900 python bench.py ) > $O/bench_b.json 2> $O/bench_b.err
( time timeout 600 python bench.py --precision fp64 --no-extras ) > $O/bench_fp64_b.json 2> $O/bench_fp64_b.err
( time timeout 600 python bench.py --precision fp64 --math parity --ticks-per-step 1500 --no-extras ) > $O/bench_parity_b.json 2> $O/bench_parity_b.err
( time timeout 600 python bench.py --hk off --no-extras ) > $O/bench_hkoff_b.json 2> $O/bench_hkoff_b.err
( time timeout 600 python bench.py --config c4 --no-extras ) > $O/bench_c4_rates_b.json 2> $O/bench_c4_rates_b.err
( time timeout 600 python bench.py --config c4 --c4-mode full --no-extras ) > $O/bench_c4_full_b.json 2> $O/bench_c4_full_b.err
AGF_NO_WARM=1 AGF_PROF_HK=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:step_kernel -s 1 -c 1 --csv --log-file $O/traffic_c3_fp32_fast_hk_131072x500.csv python profiles/prof_step.py fp32 uwb 131072 500 2 > $O/traffic_b.log 2>&1
AGF_NO_WARM=1 AGF_PROF_HK=1 AGF_PROF_C4=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:step_kernel -s 1 -c 1 --csv --log-file $O/traffic_c4_fp32_rates_2097152x64.csv python profiles/prof_step.py fp32 rates 2097152 64 2 > $O/traffic_c4_b.log 2>&1
AGF_NO_WARM=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_b.csv python bench.py --steps 3 --warmup 3 --ticks-per-step 500 --no-extras > $O/launches_b.log 2>&1
echo done
