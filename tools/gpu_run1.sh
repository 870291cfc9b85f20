#!/bin/bash
# one GPU session: tests, benches, profiles (outputs in gpurun_out/)
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r1b_pytest.log
python bench.py --workload c3 --steps 50 > gpurun_out/r1b_bench_c3.json 2> gpurun_out/r1b_bench_c3.err; tail -c 1500 gpurun_out/r1b_bench_c3.json
python bench.py --workload c3 --steps 50 --no-mip-fusion > gpurun_out/r1b_bench_c3_nofusion.json 2>/dev/null; tail -c 600 gpurun_out/r1b_bench_c3_nofusion.json
python bench.py --steps 100 --no-cpu-baseline > gpurun_out/r1b_bench_c2.json 2> gpurun_out/r1b_bench_c2.err; tail -c 1500 gpurun_out/r1b_bench_c2.json
python bench.py --steps 100 --no-cpu-baseline --no-mip-fusion > gpurun_out/r1b_bench_c2_nofusion.json 2>/dev/null; tail -c 400 gpurun_out/r1b_bench_c2_nofusion.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1b_launches_c3.csv python bench.py --workload c3 --steps 4 > gpurun_out/ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mip_fused -s 9 -c 3 -o gpurun_out/r1b_prof_mip_fused -f python bench.py --workload c3 --steps 4 > gpurun_out/ncu_c3_full.log 2>&1
ls -la gpurun_out | tail -8
