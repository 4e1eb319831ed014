#!/bin/bash
# dev aid: build tools/libbgnn_timing.so = libbgnn.so with -DBGNN_AGG_TIMING (per-phase cycle counters in gat_aggregate_tile)
cd "$(dirname "$0")/../bathymetric-gnn_b200" && mkdir -p /tmp/dbgbuild
for f in error graph_build stitch gat_forward proj_umma; do
  extra=""; [ $f = graph_build -o $f = stitch ] && extra="-fmad=false"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I ../include -I csrc -DBGNN_AGG_TIMING $BGNN_TIMING_FLAGS $extra -c csrc/$f.cu -o /tmp/dbgbuild/$f.o &
done; wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../tools/libbgnn_timing.so /tmp/dbgbuild/*.o -cudart static
