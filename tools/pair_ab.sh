# A/B of the two lighting kernels (one GPU): one row per lane (k_pbr_tex) against two rows per lane on the packed FP32 pipe (k_pbr_pair)
for cfg in "B2_PBR_PAIR=0" "B2_PBR_PAIR=1"; do
  echo "$cfg: $(env $cfg python bench.py --steps 200 --warmup 5 --no-cpu-baseline --extras 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['launch_ms'])")"
done
