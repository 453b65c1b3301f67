#!/bin/bash
# build_variants.sh name1:"-DFLAG ..." name2:"..."  ->  gpurun_out/../variants/libspamm_<name>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
       -Xcompiler -fPIC,-ffp-contract=off -shared $flags -I include -o variants/libspamm_$name.so \
       spamm_b200/csrc/spamm_b200.cu &
done
wait
ls -la variants
