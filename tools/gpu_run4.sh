#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for mf in 0 1; do
python bench.py --workload c3 --steps 50 --mip-fusion $mf > gpurun_out/r1d_bench_c3_mf$mf.json 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/r1d_bench_c3_mf$mf.json').read().strip().splitlines()[-1]); print('c3 mf$mf random ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'survey ms', d['variants']['survey_diffuse']['ms_per_step'], d['variants']['survey_diffuse']['frac'], 'launches', d['gpu_launches'])"
python bench.py --workload c1 --steps 200 --no-cpu-baseline --mip-fusion $mf > gpurun_out/r1d_bench_c1_mf$mf.json 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/r1d_bench_c1_mf$mf.json').read().strip().splitlines()[-1]); print('c1 mf$mf ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])"
python bench.py --workload c1 --steps 200 --no-cpu-baseline --mip-fusion $mf --graph 0 > gpurun_out/r1d_bench_c1_mf${mf}_nograph.json 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/r1d_bench_c1_mf${mf}_nograph.json').read().strip().splitlines()[-1]); print('c1 nograph mf$mf ms', d['ms_per_step'], 'launches', d['gpu_launches'])"
python bench.py --steps 100 --no-cpu-baseline --mip-fusion $mf > gpurun_out/r1d_bench_c2_mf$mf.json 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/r1d_bench_c2_mf$mf.json').read().strip().splitlines()[-1]); print('c2 mf$mf ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])"
done
