#!/bin/bash
# final tree: parity suite, smoke, headline bench, stand-in bench (device build with pooled scratch memory), reference arm
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log; tail -2 gpurun_out/bench.log | cut -c1-300
timeout 900 python bench.py --steps 3 --warmup 3 --workload dragon_standin_1080p_256spp --no-cold > gpurun_out/bench_dragon_standin.log 2>&1; tail -1 gpurun_out/bench_dragon_standin.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-200
