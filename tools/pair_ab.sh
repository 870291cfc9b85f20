# A/B of the two lighting kernels (one GPU): one row per lane (k_pbr_tex) against two rows per lane on the packed FP32 pipe (k_pbr_pair)
timeout 300 python -m pytest tests/test_gpu_pbr.py -x -q -m gpu 2>&1 | tail -1
for cfg in "B2_PBR_PAIR=0" "B2_PBR_PAIR=1"; do
  echo "$cfg: $(env $cfg python bench.py --steps 200 --warmup 5 --no-cpu-baseline --extras 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['launch_ms'], 'c1', d['extra']['c1']['ms_per_step'], d['extra']['c1']['lighting_ms'], 'c5', d['extra']['c5_one_gpu']['ms_per_step'])")"
done
