#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/stats.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
python scripts/profile_render.py --kernel bvh --spp 128 >> gpurun_out/stats.log 2>&1
python scripts/profile_render.py --kernel bvh --scene suzanne --spp 128 >> gpurun_out/stats.log 2>&1
grep -v "^Scene has" gpurun_out/stats.log
timeout 1500 python bench.py --steps 3 --warmup 3 --workload dragon_standin_1080p_256spp > gpurun_out/bench_dragon.log 2>&1; tail -1 gpurun_out/bench_dragon.log | cut -c1-1500
