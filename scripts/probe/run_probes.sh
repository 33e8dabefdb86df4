#!/bin/bash
# every probe in its own process (a faulting probe cannot poison the next one)
cd "$(dirname "$0")/../.."
for p in tmemst tma128 tma32 qk40 qk80 qk160 qk16 pv40 pv80 pv160 pv16 ts40 ts80 ts160 amn40 amn80; do
  timeout 60 scripts/probe/_bin/probe_tc $p 2>&1 | tail -4 || echo "PROBE $p: EXIT $?"
done
