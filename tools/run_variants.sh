#!/bin/bash
# on the GPU box: run tools/kbench.py once per variants/libspamm_*.so
cd "$(dirname "$0")/.."
for f in variants/libspamm_*.so; do
  n=$(basename $f .so); n=${n#libspamm_}
  echo "== $n"
  SPAMM_B200_LIB=$PWD/$f python tools/kbench.py 2>&1 | tail -${KB_TAIL:-1}
done
