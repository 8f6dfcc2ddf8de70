#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/stats.log
python scripts/profile_render.py --kernel bvh --spp 128 >> gpurun_out/stats.log 2>&1
python scripts/profile_render.py --kernel bvh --spp 32 --stats >> gpurun_out/stats.log 2>&1
python scripts/profile_render.py --kernel bvh --scene suzanne --spp 128 >> gpurun_out/stats.log 2>&1
python scripts/profile_render.py --kernel bvh --scene suzanne --spp 32 --stats >> gpurun_out/stats.log 2>&1
grep -v "^Scene has" gpurun_out/stats.log
