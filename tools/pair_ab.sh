# A/B of the lighting kernels (one GPU): parity of the default, then timings per variant
timeout 400 python -m pytest tests/test_gpu_pbr.py tests/test_golden.py -x -q -m gpu 2>&1 | tail -2
for cfg in "B2_PBR_PAIR=0" "B2_PAIR_CTAS=3" "B2_PAIR_CTAS=4" "B2_PAIR_CTAS=5" "B2_PAIR_CTAS=5 B2_TAIL_PCT=0" "B2_PAIR_CTAS=5 B2_TAIL_PCT=15"; do
  echo "$cfg: $(env $cfg python bench.py --steps 200 --warmup 5 --no-cpu-baseline --extras 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['launch_ms'])")"
done
