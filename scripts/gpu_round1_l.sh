#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/stats.log
python scripts/check_variant.py 201 202 205 208 2>&1 | grep -v "^Scene has" | tee gpurun_out/check_variant.log
for v in 0 201 202 203 204 205 206 207 208 209; do
  python scripts/profile_render.py --kernel bvh --rays-per-lane $v --spp 128 >> gpurun_out/stats.log 2>&1
done
for v in 0 202 208; do
  python scripts/profile_render.py --kernel bvh --rays-per-lane $v --scene suzanne --spp 128 >> gpurun_out/stats.log 2>&1
done
grep -v "^Scene has" gpurun_out/stats.log
