#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_mip_fused.py tests/test_gpu_mip.py tests/test_gpu_pbr.py -x -q -m gpu 2>&1 | tail -4
python bench.py --workload c3 --steps 50 > gpurun_out/r1c_bench_c3.json 2>&1; tail -c 1300 gpurun_out/r1c_bench_c3.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -s 9 -c 6 --csv --log-file gpurun_out/r1c_launches_c3_random.csv python bench.py --workload c3 --steps 4 --c3-data random > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -s 9 -c 6 --csv --log-file gpurun_out/r1c_launches_c3_survey.csv python bench.py --workload c3 --steps 4 --c3-data survey > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mip_fused -s 9 -c 1 -o gpurun_out/r1c_prof_mip_fused_survey -f python bench.py --workload c3 --steps 4 --c3-data survey > /dev/null 2>&1
ls -la gpurun_out | tail -5
