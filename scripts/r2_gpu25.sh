#!/bin/bash
# K2w loop shape: V inner-node visits per leaf check, R repetitions per traversal phase
P="python scripts/profile_render.py"
for rep in 1 2; do
for tag in product v3r6 v3r7 v4r5 v4r6 v5r4 v6r3; do
  lib=raytracing-one-weekend_b200/librtw_b200_$tag.so; [ $tag = product ] && lib=""
  RTW_LIB=$lib $P --kernel bvh --spp 128 2>&1 | tail -1 | sed "s/^/$tag /"
done; done
