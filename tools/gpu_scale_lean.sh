#!/bin/bash
# lean 8-GPU check: C5 at N = 1 and 8, default workload at N = 8
mkdir -p gpurun_out
./tools/gpu_c5.sh 1 8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29608 bench.py --gpus 8 --steps 100 --warmup 3 > gpurun_out/r1j_c2_n8.json 2> gpurun_out/r1j_c2_n8.err
python -c "
import json; d=json.loads(open('gpurun_out/r1j_c2_n8.json').read().strip().splitlines()[-1]); print('c2 N=8 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])" || tail -3 gpurun_out/r1j_c2_n8.err
