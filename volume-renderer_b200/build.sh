#!/usr/bin/env bash
# Builds libvolren_b200.so (CUDA kernels + C-ABI) and libvolren_host.so (C++ host mirror of the
# reference's RendererCore/Camera/loader surface) in-tree for sm_100a.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="$HERE/lib"
mkdir -p "$OUT"
NVFLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false
         -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden,-Wall
         -Xptxas -v -I"$ROOT/include" -I"$HERE/csrc")
echo "[build] nvcc volren_abi.cu -> lib/libvolren_b200.so"
"$NVCC" "${NVFLAGS[@]}" -shared -o "$OUT/libvolren_b200.so" "$HERE/csrc/volren_abi.cu" 2> "$OUT/ptxas.log" || { cat "$OUT/ptxas.log"; exit 1; }
grep -E "error|warning" "$OUT/ptxas.log" | grep -v "ptxas info" || true
if ls "$HERE"/host/*.cpp >/dev/null 2>&1; then
  echo "[build] g++ host/*.cpp -> lib/libvolren_host.so"
  g++ -O2 -std=c++17 -fPIC -ffp-contract=off -fvisibility=hidden -Wall -shared \
      -I"$ROOT/include" -I"$HERE/host" -o "$OUT/libvolren_host.so" "$HERE"/host/*.cpp \
      -L"$OUT" -lvolren_b200 -Wl,-rpath,'$ORIGIN'
fi
echo "[build] done"
