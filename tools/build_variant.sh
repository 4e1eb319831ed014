#!/bin/bash
# dev aid: build tools/libbgnn_<tag>.so with extra nvcc flags:  tools/build_variant.sh <tag> [-DFOO=1 ...]
tag=$1; shift
cd "$(dirname "$0")/../bathymetric-gnn_b200" && mkdir -p /tmp/varbuild_$tag
for f in error graph_build stitch vr_post gat_forward proj_umma proj_bf16; do
  extra=""; [ $f = graph_build -o $f = stitch -o $f = vr_post ] && extra="-fmad=false"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I ../include -I csrc "$@" $extra -c csrc/$f.cu -o /tmp/varbuild_$tag/$f.o &
done; wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../tools/libbgnn_$tag.so /tmp/varbuild_$tag/*.o -cudart static
