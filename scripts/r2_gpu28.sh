#!/bin/bash
# traversal conditions without the lane state (product) against the previous commit's build; pooled scratch memory of the device BVH build
python scripts/e2e_probe_mesh2.py 2>&1 | grep -v "^Scene"
P="python scripts/profile_render.py"
for rep in 1 2; do
for tag in product prev; do
  lib=raytracing-one-weekend_b200/librtw_b200_$tag.so; [ $tag = product ] && lib=""
  RTW_LIB=$lib $P --kernel bvh --spp 128 2>&1 | tail -1 | sed "s/^/$tag /"
  RTW_LIB=$lib $P --kernel bvh --scene suzanne --spp 64 --depth 20 2>&1 | tail -1 | sed "s/^/$tag /"
  RTW_LIB=$lib $P --kernel bvh --scene /tmp/standin5.obj --spp 64 --depth 20 --bvh-build host 2>&1 | tail -1 | sed "s/^/$tag /"
done; done
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
