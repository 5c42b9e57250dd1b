#!/usr/bin/env bash
# builds tools/marchlab (development kernel lab) for sm_100a
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -lineinfo $LABFLAGS -I../include -I../volume-renderer_b200/csrc marchlab.cu -o marchlab
