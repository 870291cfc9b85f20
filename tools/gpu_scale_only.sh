#!/bin/bash
# strong-scaling sweep of the default multi-GPU workload (C5 row bands) on one 8-GPU box: N = 8, 4, 2 (no host-bandwidth probe)
T=${1:-r2}
mkdir -p gpurun_out
for N in 8 4 2; do
  B2_BAND_DEBUG=1 timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 50 --warmup 3 --extras 0 > gpurun_out/${T}_scale_n$N.json 2> gpurun_out/${T}_scale_n$N.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${T}_scale_n$N.json"))
print($N, d["ms_per_step"], d["sustained"]["ms_per_step"], d["parity"]["bands_equal_unbanded_frame"], d["strong_scaling_in_this_run"], "e2e", d["e2e"]["ms_per_frame"])
PY
done
