#!/bin/bash
# Kernel-tuning builds: compiles ONE source with -D overrides into aidet_b200/libexp_<tag>.so (the other objects come
# from the regular build), to be selected with AIDET_B200_LIB=$PWD/aidet_b200/libexp_<tag>.so, e.g.
#   scripts/build_variants.sh riou dual0 -DAIDET_RIOU_DUAL=0 -DAIDET_RIOU_MINB=1
#   scripts/build_variants.sh rnms occ3  -DAIDET_NMS_MINB=3
#   AIDET_B200_LIB=$PWD/aidet_b200/libexp_riou_dual0.so python scripts/riou_time.py
# Knobs: AIDET_RIOU_MINB / AIDET_RIOU_UNROLL / AIDET_RIOU_DUAL (csrc/riou.cu), AIDET_NMS_MINB (csrc/rnms.cu).
# The variant libraries are git-ignored (*.so); delete them before a gpurun call that does not need them (they travel).
set -euo pipefail
src=$1; tag=$2; shift 2
here=$(cd "$(dirname "$0")/.." && pwd)
cd "$here/aidet_b200/csrc"
make -j4 >/dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr \
     -ccbin /usr/bin/g++ "$@" -c "$src.cu" -o "/tmp/aidet_exp_${src}_${tag}.o" 2> "/tmp/aidet_exp_${src}_${tag}.log"
grep -E "Used [0-9]+ registers" "/tmp/aidet_exp_${src}_${tag}.log" | sort | uniq -c | sort -rn | head -5
objs=""
for o in capi riou riou_assign rnms rroi_align soft_nms; do
  if [ "$o" = "$src" ]; then objs="$objs /tmp/aidet_exp_${src}_${tag}.o"; else objs="$objs $o.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "../libexp_${src}_${tag}.so" $objs -lcudart -ccbin /usr/bin/g++
echo "built aidet_b200/libexp_${src}_${tag}.so"
