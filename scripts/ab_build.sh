#!/bin/bash
# usage: scripts/ab_build.sh name "-DFLAG=.. -DFLAG2=.." [name2 "flags2" ...]  — A/B builds of librtnw.so into ab/ (tuning aid; select with RTNW_LIB)
ROOT=$(cd "$(dirname "$0")/.." && pwd)
PKG="$ROOT/peter-shirley-ray-tracing-the-next-week_b200"
mkdir -p "$ROOT/ab"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( /usr/local/cuda/bin/nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I"$ROOT/include" \
      --fmad=false -Xptxas -v --expt-relaxed-constexpr -DRTNW_TUNING $flags -shared -o "$ROOT/ab/$name.so" "$PKG/csrc/rtnw_cuda.cu" > "$ROOT/ab/$name.log" 2>&1 \
      || { echo "BUILD FAILED $name"; grep -E "error" "$ROOT/ab/$name.log" | head; }
    echo "$name: $(grep -A2 'k_renderILb0' "$ROOT/ab/$name.log" | grep -E 'Used|spill' | tr '\n' ' ' | sed 's/ptxas info *: //;s/  */ /g')" ) &
done
wait
