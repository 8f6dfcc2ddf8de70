#!/bin/bash
# K2w loop shape: {visit, leaf} x 16 (product) against {visit, visit, leaf} x 8 and the 2x-unrolled {visit, leaf, visit, leaf} x 8
P="python scripts/profile_render.py"
for rep in 1 2; do
for tag in product v2 v3; do
  lib=raytracing-one-weekend_b200/librtw_b200_$tag.so; [ $tag = product ] && lib=""
  RTW_LIB=$lib $P --kernel bvh --spp 128 2>&1 | tail -1 | sed "s/^/$tag /"
done; done
RTW_LIB=raytracing-one-weekend_b200/librtw_b200_v2.so $P --kernel bvh --spp 16 --stats 2>&1 | tail -1 | sed "s/^/v2 /"
