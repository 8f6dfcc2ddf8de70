#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/stats.log
for v in 0 101 102 103 104 105; do
  python scripts/profile_render.py --kernel bvh --rays-per-lane $v --spp 128 >> gpurun_out/stats.log 2>&1
done
for leaf in 1 2 3 6 8; do
  echo "leaf=$leaf" >> gpurun_out/stats.log
  RTW_BVH_LEAF=$leaf python scripts/profile_render.py --kernel bvh --rays-per-lane 101 --spp 128 >> gpurun_out/stats.log 2>&1
  RTW_BVH_LEAF=$leaf python scripts/profile_render.py --kernel bvh --rays-per-lane 0 --spp 128 >> gpurun_out/stats.log 2>&1
done
echo "suzanne" >> gpurun_out/stats.log
for leaf in 2 4 8; do
  RTW_BVH_LEAF=$leaf python scripts/profile_render.py --kernel bvh --rays-per-lane 101 --scene suzanne --spp 128 >> gpurun_out/stats.log 2>&1
  RTW_BVH_LEAF=$leaf python scripts/profile_render.py --kernel bvh --rays-per-lane 0 --scene suzanne --spp 128 >> gpurun_out/stats.log 2>&1
done
cat gpurun_out/stats.log
