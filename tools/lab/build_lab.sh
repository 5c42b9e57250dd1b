#!/usr/bin/env bash
# Development build of the library with the pipeline-depth / occupancy variants of the march kernels compiled in
# (-DVR_LAB; selected per launch with VR_LAB_TP="depth,minb"): volume-renderer_b200/lib/libvolren_b200_lab.so.
# Use it through VOLREN_B200_LIB=<that file> (tools/lab/variants.py does).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$(dirname "$HERE")")"
PKG="$ROOT/volume-renderer_b200"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false -DVR_LAB ${LABFLAGS:-} \
     -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden -Xptxas -v -I"$ROOT/include" -I"$PKG/csrc" \
     -shared -o "$PKG/lib/libvolren_b200_lab.so" "$PKG/csrc/volren_abi.cu" 2> "$PKG/lib/ptxas_lab.log"
echo "built $PKG/lib/libvolren_b200_lab.so"
