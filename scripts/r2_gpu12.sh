#!/bin/bash
# K2w at 32 warps per SM (64 registers, lean lane state): record count / batch size variants, scalar slab build
P="python scripts/profile_render.py"
export RTW_LIB=raytracing-one-weekend_b200/librtw_b200_scalar.so
for rep in 1 2; do
$P --kernel bvh --spp 128 2>&1 | tail -1 | sed "s/^/28 warps P=96 B=32 /"
for v in 0 1 2 3; do RTW_WF_WARPS=32 RTW_WF_VARIANT=$v $P --kernel bvh --spp 128 2>&1 | tail -1 | sed "s/^/32 warps variant $v /"; done
done
