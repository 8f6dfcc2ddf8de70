#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
for v in "--kernel spheres --rays-per-lane 4" "--kernel spheres --rays-per-lane 2" "--kernel bvh"; do
  python scripts/profile_render.py $v --spp 32 --stats >> gpurun_out/stats.log 2>&1
  python scripts/profile_render.py $v --spp 32 >> gpurun_out/stats.log 2>&1
done
python scripts/profile_render.py --kernel bvh --scene suzanne --spp 32 --stats >> gpurun_out/stats.log 2>&1
cat gpurun_out/stats.log
# ncu: one full capture per kernel variant (second launch of each process)
python scripts/profile_render.py --kernel spheres --rays-per-lane 4 --spp 8 > gpurun_out/plain_k1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_render -s 1 -c 1 -o gpurun_out/prof_k1_r4 python scripts/profile_render.py --kernel spheres --rays-per-lane 4 --spp 8 > gpurun_out/ncu_k1.log 2>&1
python scripts/profile_render.py --kernel bvh --spp 8 > gpurun_out/plain_k2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_render -s 1 -c 1 -o gpurun_out/prof_k2 python scripts/profile_render.py --kernel bvh --spp 8 > gpurun_out/ncu_k2.log 2>&1
tail -3 gpurun_out/ncu_k1.log gpurun_out/ncu_k2.log
ls -la gpurun_out
