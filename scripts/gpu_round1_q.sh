#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/stats.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
python scripts/profile_render.py --kernel bvh --spp 128 >> gpurun_out/stats.log 2>&1
grep -v "^Scene has" gpurun_out/stats.log
timeout 1500 python bench.py --steps 2 --warmup 3 --workload dragon_standin_1080p_256spp --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'])"
python -c "import __graft_entry__ as g; g.smoke()"
