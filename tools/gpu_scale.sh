#!/bin/bash
# multi-GPU session on one 8-GPU box: C5 (16384^2 canvas, row bands, strong scaling) at N = 1, 2, 4, 8 and the default
# workload (independent frames, weak scaling) at N = 8
mkdir -p gpurun_out
nvidia-smi -L | head -8
python bench.py --workload c5 --steps 10 > gpurun_out/r1g_c5_n1.json 2> gpurun_out/r1g_c5_n1.err; tail -c 700 gpurun_out/r1g_c5_n1.json
for N in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --workload c5 --steps 10 > gpurun_out/r1g_c5_n$N.json 2> gpurun_out/r1g_c5_n$N.err
  tail -c 700 gpurun_out/r1g_c5_n$N.json; tail -3 gpurun_out/r1g_c5_n$N.err | cut -c1-300
done
for N in 2 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 100 --warmup 3 > gpurun_out/r1g_c2_n$N.json 2> gpurun_out/r1g_c2_n$N.err
tail -c 900 gpurun_out/r1g_c2_n$N.json; tail -3 gpurun_out/r1g_c2_n$N.err | cut -c1-300
done
