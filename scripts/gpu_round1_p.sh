#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/stats.log
python scripts/profile_render.py --kernel bvh --spp 128 >> gpurun_out/stats.log 2>&1
python scripts/profile_render.py --kernel bvh --spp 128 --rays-per-lane 1000 >> gpurun_out/stats.log 2>&1
python scripts/profile_render.py --kernel bvh --scene suzanne --spp 128 >> gpurun_out/stats.log 2>&1
grep -v "^Scene has" gpurun_out/stats.log
timeout 1500 python bench.py --steps 2 --warmup 3 --workload dragon_standin_1080p_256spp --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'])"
