#!/bin/bash
# K2w node step without branches (predicated push / pop): A/B on one box against the branchy build, scalar and FFMA2 slab tests
mkdir -p gpurun_out
P="python scripts/profile_render.py"
for rep in 1 2; do
for tag in product scalar bf bf2; do
  lib=raytracing-one-weekend_b200/librtw_b200_$tag.so; [ $tag = product ] && lib=""
  RTW_LIB=$lib $P --kernel bvh --spp 128 2>&1 | tail -1 | sed "s/^/$tag /"
done; done
RTW_LIB=raytracing-one-weekend_b200/librtw_b200_bf.so $P --kernel bvh --spp 1024 2>&1 | tail -1 | sed "s/^/bf /"
