#!/bin/bash
# K2w node visit: without the (unreachable) stack guard; two-way branch (some child hit / none) with a conditional push
P="python scripts/profile_render.py"
for rep in 1 2; do
for tag in product NOGUARD ANYNONE; do
  lib=raytracing-one-weekend_b200/librtw_b200_$tag.so; [ $tag = product ] && lib=""
  RTW_LIB=$lib $P --kernel bvh --spp 128 2>&1 | tail -1 | sed "s/^/$tag /"
done; done
