#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/stats.log
for v in 101 102 103 104 105 100; do
  python scripts/profile_render.py --kernel bvh --rays-per-lane $v --spp 128 >> gpurun_out/stats.log 2>&1
done
cat gpurun_out/stats.log
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -3 gpurun_out/bench.log
