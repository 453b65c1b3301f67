#!/bin/bash
# tools/build_variant.sh NAME [-DMACRO ...]: kernel-iteration build of the library (config-5 instantiation
# only) into variants/libNAME.so; run with SPAMM_B200_LIB=variants/libNAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
     -Xcompiler -fPIC,-ffp-contract=off -shared -DSPAMM_DEV_BUILD "$@" -I include \
     -o variants/lib$name.so spamm_b200/csrc/spamm_b200.cu
echo built variants/lib$name.so
