#!/bin/bash
python -m pytest tests/test_gpu_mip_fused.py tests/test_gpu_mip.py -x -q -m gpu 2>&1 | tail -2
python bench.py --workload c3 --steps 50 > gpurun_out/r1i_bench_c3.json 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/r1i_bench_c3.json').read().strip().splitlines()[-1]); print('c3 random ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'survey ms', d['variants']['survey_diffuse']['ms_per_step'], d['variants']['survey_diffuse']['frac'], 'launches', d['gpu_launches'])"
ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 5 --csv --log-file gpurun_out/r1i_launches_c3.csv python bench.py --workload c3 --steps 4 --c3-data survey > /dev/null 2>&1
grep -v "^==" gpurun_out/r1i_launches_c3.csv | python -c "
import csv,sys
for r in list(csv.reader(sys.stdin))[1:]: print(r[4][:50], r[8], r[-1])
"
