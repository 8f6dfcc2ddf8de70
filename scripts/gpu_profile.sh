#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command, full captures of the dominant kernels (each only after the same
# command has exited 0 without ncu).   gpurun --timeout 1500 -- 'bash scripts/gpu_profile.sh'
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --spp 64 --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv $B > gpurun_out/ncu_launch.log 2>&1
prof() {  # name, then the arguments of scripts/profile_render.py
  local name=$1; shift
  python scripts/profile_render.py "$@" > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_render -s 1 -c 1 -f -o gpurun_out/prof_$name python scripts/profile_render.py "$@" > gpurun_out/ncu_$name.log 2>&1
}
prof k2w --kernel bvh --spp 8
prof k2_perlane --kernel bvh-perlane --spp 8
prof k1 --kernel spheres --spp 8
python -c "
import importlib,ctypes as C,sys
sys.path.insert(0,'.')
rtw=importlib.import_module('raytracing-one-weekend_b200')
n=C.c_longlong(0); rtw.host().rtwh_make_mesh(b'assets/suzanne.obj', b'/tmp/standin5.obj', 5, 20221018, 0.08, C.byref(n)); print('tris', n.value)
"
prof k2_dragon --kernel bvh --scene /tmp/standin5.obj --spp 4 --depth 20
prof k2_suzanne --kernel bvh --scene suzanne --spp 8 --depth 20
cat gpurun_out/plain_k2w.log gpurun_out/plain_k2_perlane.log gpurun_out/plain_k1.log gpurun_out/plain_k2_dragon.log gpurun_out/plain_k2_suzanne.log | grep -v "^Scene"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_bench.csv
