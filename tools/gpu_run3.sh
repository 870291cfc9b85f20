#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_mip_fused.py tests/test_gpu_mip.py -x -q -m gpu 2>&1 | tail -3
python bench.py --workload c3 --steps 50 > gpurun_out/r1c_bench_c3.json 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/r1c_bench_c3.json').read().strip().splitlines()[-1]); print('random ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'survey ms', d['variants']['survey_diffuse']['ms_per_step'], d['variants']['survey_diffuse']['frac'])"
for f in random survey; do
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -s 9 -c 3 --csv --log-file gpurun_out/r1c_launches_c3_$f.csv python bench.py --workload c3 --steps 4 --c3-data $f > /dev/null 2>&1
grep -v "^==" gpurun_out/r1c_launches_c3_$f.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))[1:]
from collections import OrderedDict
d=OrderedDict()
for r in rows:
    d.setdefault(r[0],[r[4][27:56],r[8]]).append((r[-3].split('__')[1][:14],r[-1]))
for k,v in list(d.items())[:3]: print(k,v)
"
done
