#!/bin/bash
# C5 on N GPUs of this box: ./tools/gpu_c5.sh 2 4 8
mkdir -p gpurun_out
for N in "$@"; do
  if [ "$N" = "1" ]; then python bench.py --workload c5 --steps 10 > gpurun_out/r1j_c5_n1.json 2> gpurun_out/r1j_c5_n1.err
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --workload c5 --steps 10 > gpurun_out/r1j_c5_n$N.json 2> gpurun_out/r1j_c5_n$N.err; fi
  python -c "
import json; d=json.loads(open('gpurun_out/r1j_c5_n$N.json').read().strip().splitlines()[-1]); print('c5 N=$N ms', d['ms_per_step'], 'value', d['value'])" || tail -5 gpurun_out/r1j_c5_n$N.err
done
