#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, ONE GPU). Numbers printed by runs under ncu are never bench values.
set -u
OUT=gpurun_out
mkdir -p $OUT
TAG=${1:-r01}
# 1. every launch of a reduced-step bench run with its device time (cold-cache, serialised: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches_bench.csv \
  python bench.py --steps 1 --warmup 0 --ddim-steps 2 --no-cpu-baseline > $OUT/${TAG}_launches_bench.stdout 2> $OUT/${TAG}_launches_bench.stderr
echo "launch list rc=$?"
# 2. full captures of the hand-written kernels (first launches of the microbenchmarks: 16 x 64 x 64 x 320 layer shapes)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:temporal_attn_fwd_kernel -c 2 -f -o $OUT/${TAG}_temporal_fwd \
  python scripts/kernel_bench.py --ncu > /dev/null 2>&1
echo "temporal fwd rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:temporal_attn_bwd_kernel -c 1 -f -o $OUT/${TAG}_temporal_bwd \
  python scripts/kernel_bench.py --ncu > /dev/null 2>&1
echo "temporal bwd rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cross_attn -c 1 -f -o $OUT/${TAG}_cross_attn_fwd \
  python scripts/xattn_bench.py > /dev/null 2>&1
echo "cross fwd rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cross_attn_bwd -c 1 -f -o $OUT/${TAG}_cross_attn_bwd \
  python scripts/xattn_bench.py > /dev/null 2>&1
echo "cross bwd rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:groupnorm -c 2 -f -o $OUT/${TAG}_groupnorm_fwd \
  python scripts/glue_bench.py > /dev/null 2>&1
echo "groupnorm rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:geglu_lut -c 1 -f -o $OUT/${TAG}_geglu_lut \
  python scripts/glue_bench.py > /dev/null 2>&1
echo "geglu rc=$?"
ls -la $OUT | tail -n 15
