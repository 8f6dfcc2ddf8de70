#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/stats.log
for v in 101 102 103 104 105 106 107 108 109; do
  python scripts/profile_render.py --kernel bvh --rays-per-lane $v --spp 128 >> gpurun_out/stats.log 2>&1
done
cat gpurun_out/stats.log
