#!/bin/bash
# (1) launch list of one guided + one plain DDIM step with per-launch device time; (2) DRAM bytes of every temporal_attn_fwd
# launch of the same two steps (roofline.traffic). One GPU; numbers under ncu are never bench values.
set -u
OUT=gpurun_out; TAG=${1:-r02}
mkdir -p $OUT
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $OUT/${TAG}_launches_step.csv \
  python scripts/launch_list_step.py > $OUT/${TAG}_launches_step.stdout 2>&1
echo "launch list rc=$?"
gzip -f $OUT/${TAG}_launches_step.csv
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k regex:temporal_attn_fwd_kernel -c 400 --csv --log-file $OUT/${TAG}_temporal_traffic.csv \
  python scripts/launch_list_step.py > $OUT/${TAG}_temporal_traffic.stdout 2>&1
echo "traffic rc=$?"
gzip -f $OUT/${TAG}_temporal_traffic.csv
ls -la $OUT | grep ${TAG}_
