#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/stats.log
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
for v in 100 101 102 103 104 105 106 107 108 1102; do
  python scripts/profile_render.py --kernel bvh --rays-per-lane $v --spp 128 >> gpurun_out/stats.log 2>&1
done
python scripts/profile_render.py --kernel bvh --rays-per-lane 102 --scene suzanne --spp 64 >> gpurun_out/stats.log 2>&1
python scripts/profile_render.py --kernel bvh --rays-per-lane 100 --scene suzanne --spp 64 >> gpurun_out/stats.log 2>&1
cat gpurun_out/stats.log
python scripts/profile_render.py --kernel bvh --rays-per-lane 102 --spp 8 > gpurun_out/plain_k2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_render -s 1 -c 1 -o gpurun_out/prof_k2_v102 python scripts/profile_render.py --kernel bvh --rays-per-lane 102 --spp 8 > gpurun_out/ncu_k2.log 2>&1
tail -n 2 gpurun_out/ncu_k2.log
