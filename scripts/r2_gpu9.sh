#!/bin/bash
# full ncu captures of K2w and the per-lane kernel with the FFMA2 node test
mkdir -p gpurun_out
prof() {  # name, then the arguments of scripts/profile_render.py
  local name=$1; shift
  python scripts/profile_render.py "$@" > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_render -s 1 -c 1 -f -o gpurun_out/prof_$name python scripts/profile_render.py "$@" > gpurun_out/ncu_$name.log 2>&1
}
prof k2w_f2 --kernel bvh --spp 8
prof k2_perlane_f2 --kernel bvh-perlane --spp 8
cat gpurun_out/plain_k2w_f2.log gpurun_out/plain_k2_perlane_f2.log
