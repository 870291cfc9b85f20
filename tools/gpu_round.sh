#!/bin/bash
# end-of-milestone GPU session: tests, bench lines, launch list, one full capture of the top kernels
T=${1:-r1f}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.log
python bench.py > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; tail -c 2200 gpurun_out/${T}_bench_c2.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>/dev/null; tail -c 500 gpurun_out/${T}_bench_reference.json
python bench.py --workload c1 --steps 300 > gpurun_out/${T}_bench_c1.json 2>/dev/null; tail -c 900 gpurun_out/${T}_bench_c1.json
python bench.py --workload c3 --steps 50 > gpurun_out/${T}_bench_c3.json 2>/dev/null; tail -c 700 gpurun_out/${T}_bench_c3.json
python bench.py --workload c4 --steps 5 --no-cpu-baseline > gpurun_out/${T}_bench_c4.json 2>/dev/null; tail -c 600 gpurun_out/${T}_bench_c4.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --graph 0 > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pbr_fast -s 4 -c 1 -o gpurun_out/${T}_prof_pbr_fast -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --graph 0 > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out | grep ${T}
