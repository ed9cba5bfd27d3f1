#!/usr/bin/env bash
# Builds simuverse_b200/_native/liblbm_b200.so for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../_native"
mkdir -p "$out"
name="${LBM_OUT_NAME:-liblbm_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
HOSTCXX="${LBM_HOSTCXX:-/usr/bin/g++}"
[ -x "$HOSTCXX" ] || HOSTCXX=g++
set -x
"$NVCC" -ccbin "$HOSTCXX" -std=c++17 -O3 -lineinfo \
  -gencode arch=compute_100a,code=sm_100a \
  -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  -Xptxas -v -Xcompiler -fPIC,-O2,-ffp-contract=off,-Wall \
  -shared -cudart static \
  -o "$out/$name" "$here/lbm_b200.cu" "$here/host_logic.cpp" ${LBM_EXTRA_NVCC_FLAGS:-}
