#!/bin/bash
# build_variant.sh NAME "-DFOO=1 ..." file.cu [file.cu ...]: a side-by-side library for A/B runs (python scripts/sattn_bench.py --lib motionclone_b200/libmc_variant_NAME.so):
# the named sources are recompiled with the extra defines, every other object comes from the regular build.
set -e
cd "$(dirname "$0")/.."
name=$1; defs=$2; shift 2
B=motionclone_b200/csrc/_build; V=$B/variant_$name; mkdir -p $V
objs=""
for o in $B/*.o; do
  base=$(basename $o .o); skip=0
  for f in "$@"; do [ "$base" == "${f%.cu}" ] && skip=1; done
  [ $skip == 0 ] && objs="$objs $o"
done
for f in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $defs -c motionclone_b200/csrc/$f -o $V/${f%.cu}.o &
done
wait
for f in "$@"; do objs="$objs $V/${f%.cu}.o"; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o motionclone_b200/libmc_variant_$name.so $objs
echo built motionclone_b200/libmc_variant_$name.so
