#!/bin/bash
mkdir -p gpurun_out
python scripts/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1; cat gpurun_out/e2e_probe.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log; tail -2 gpurun_out/bench.log
timeout 900 python bench.py --steps 3 --warmup 3 --workload suzanne_on_ground_1080p_256spp > gpurun_out/bench_suzanne.log 2>&1; tail -1 gpurun_out/bench_suzanne.log
timeout 1500 python bench.py --steps 3 --warmup 3 --workload dragon_standin_1080p_256spp > gpurun_out/bench_dragon.log 2>&1; tail -1 gpurun_out/bench_dragon.log
timeout 900 python bench.py --steps 3 --warmup 3 --workload cover_default_200x133_20spp_depth20 > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log
