#!/bin/bash
# One GPU: every bench workload (BASELINE.json configs 1-4).   gpurun --timeout 3000 -- 'bash scripts/gpu_workloads.sh'
mkdir -p gpurun_out
for w in cover_default_200x133_20spp_depth20 cover_1080p_1024spp_depth50 suzanne_on_ground_1080p_256spp dragon_standin_1080p_256spp; do
  timeout 1500 python bench.py --steps 3 --warmup 3 --workload $w > gpurun_out/bench_$w.log 2>&1; tail -1 gpurun_out/bench_$w.log | cut -c1-400
done
