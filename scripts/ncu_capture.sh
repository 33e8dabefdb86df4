#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, ONE GPU). Numbers printed by runs under ncu are never bench values.
# Reports are exported to CSV on the box and the .ncu-rep files dropped (gpurun_out/ is capped at 64 MiB).
set -u
OUT=gpurun_out
mkdir -p $OUT
TAG=${1:-r01}
full() {  # name, kernel regex, count, command...
  local name=$1 rx=$2 cnt=$3; shift 3
  timeout 240 ncu --set full --clock-control none -k regex:$rx -c $cnt -f -o /tmp/${name} "$@" > /dev/null 2>&1
  echo "$name rc=$?"
  ncu -i /tmp/${name}.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/${name}.ncu-rep --page details --csv > $OUT/${TAG}_ncu_${name}_details.csv 2>/dev/null
}
full temporal_fwd temporal_attn_fwd_kernel 2 python scripts/kernel_bench.py --ncu
full temporal_bwd temporal_attn_bwd_kernel 1 python scripts/kernel_bench.py --ncu
full cross_attn_fwd cross_attn_fwd 1 python scripts/xattn_bench.py
full cross_attn_bwd cross_attn_bwd 1 python scripts/xattn_bench.py
full groupnorm_fwd groupnorm 2 python scripts/glue_bench.py
if [ "${2:-}" = "launches" ]; then
  # every launch of one guided + one plain DDIM step with its device time (cold-cache, serialised: compare SHARES)
  timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $OUT/${TAG}_launches_step.csv \
    python scripts/launch_list_step.py > $OUT/${TAG}_launches_step.stdout 2>&1
  echo "launch list rc=$?"
  gzip -f $OUT/${TAG}_launches_step.csv
fi
du -sh $OUT; ls -la $OUT | tail -n 14
