#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command, full captures of the dominant kernels.
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --spp 64 --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv $B > gpurun_out/ncu_launch.log 2>&1
python scripts/profile_render.py --kernel bvh --spp 8 > gpurun_out/plain_k2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_render -s 1 -c 1 -o gpurun_out/prof_k2_final python scripts/profile_render.py --kernel bvh --spp 8 > gpurun_out/ncu_k2.log 2>&1
python scripts/profile_render.py --kernel spheres --spp 8 > gpurun_out/plain_k1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_render -s 1 -c 1 -o gpurun_out/prof_k1_final python scripts/profile_render.py --kernel spheres --spp 8 > gpurun_out/ncu_k1.log 2>&1
python -c "
import importlib,ctypes as C,sys
sys.path.insert(0,'.')
rtw=importlib.import_module('raytracing-one-weekend_b200')
n=C.c_longlong(0); rtw.host().rtwh_make_mesh(b'tests/golden/suzanne.obj', b'/tmp/standin5.obj', 5, 20221018, 0.08, C.byref(n)); print('tris', n.value)
"
python scripts/profile_render.py --kernel bvh --scene /tmp/standin5.obj --spp 4 --depth 20 > gpurun_out/plain_dragon.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_render -s 1 -c 1 -o gpurun_out/prof_k2_dragon python scripts/profile_render.py --kernel bvh --scene /tmp/standin5.obj --spp 4 --depth 20 > gpurun_out/ncu_dragon.log 2>&1
cat gpurun_out/plain_k2.log gpurun_out/plain_k1.log gpurun_out/plain_dragon.log | grep -v "^Scene"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_bench.csv
