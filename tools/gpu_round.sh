#!/bin/bash
# end-of-milestone GPU session on ONE GPU: tests, smoke, the default bench line, the reference arm, the launch list
T=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; tail -c 300 gpurun_out/${T}_bench_c2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/${T}_bench_reference.json
B2_PIPE_DEBUG=1 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --extras 0 2>&1 >/dev/null | grep "b2pbr pipe" | tail -6
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --graph 0 --extras 0 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_frame.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --graph 0 --extras 0 > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pbr_pair -s 4 -c 1 -o gpurun_out/${T}_prof_pbr_pair -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --graph 0 --extras 0 > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out | grep ${T}_
