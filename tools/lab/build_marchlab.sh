#!/usr/bin/env bash
# builds tools/lab/marchlab: the round-1 development kernels (LSU, two-gather, two-ray, hybrid, TMA-windowed) on the
# headline workload, each checked bit for bit against the generic kernel.  Development tool, not product.
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -lineinfo ${LABFLAGS:-} -I../../include -I../../volume-renderer_b200/csrc -I. marchlab.cu -o marchlab
