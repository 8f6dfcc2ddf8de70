#!/bin/bash
# First GPU pass: parity tests, kernel-variant timing at reduced spp, the headline bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for v in "--kernel spheres --rays-per-lane 4" "--kernel spheres --rays-per-lane 2" "--kernel spheres --rays-per-lane 1" "--kernel bvh"; do
  echo "== $v" >> gpurun_out/variants.log
  timeout 600 python bench.py --steps 2 --warmup 3 --spp 128 --no-cpu-baseline --no-e2e $v >> gpurun_out/variants.log 2>&1
done
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -5 gpurun_out/smoke.log; tail -15 gpurun_out/pytest_gpu.log; cat gpurun_out/variants.log; tail -3 gpurun_out/bench.log
