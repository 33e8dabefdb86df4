#!/bin/bash
# ncu --set full captures of the spatial self-attention kernels (one GPU). CSV exports land in gpurun_out/.
set -u
OUT=gpurun_out; TAG=${1:-r02}; SHAPE=${2:-self64}
mkdir -p $OUT
cap() {  # name, kernel regex, skip, count
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o /tmp/$1 python scripts/sattn_bench.py $SHAPE > /dev/null 2>&1
  echo "$1 rc=$?"
  ncu -i /tmp/$1.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details --csv > $OUT/${TAG}_ncu_$1_details.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv > $OUT/${TAG}_ncu_$1_source.csv 2>/dev/null
  gzip -f $OUT/${TAG}_ncu_$1_source.csv
}
cap sattn_fwd spatial_attn_fwd_kernel 1 1
cap sattn_bwd_dq spatial_attn_bwd_dq_kernel 1 1
cap sattn_bwd_dkv spatial_attn_bwd_dkv_kernel 1 1
ls -la $OUT | tail -12
