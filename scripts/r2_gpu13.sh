#!/bin/bash
# 32-warp lean tier of K2w as the default: full parity suite, speed, traversal steps per phase at 32 warps
mkdir -p gpurun_out
python -c "
import importlib,ctypes as C,sys
sys.path.insert(0,'.')
rtw=importlib.import_module('raytracing-one-weekend_b200')
n=C.c_longlong(0); rtw.host().rtwh_make_mesh(b'assets/suzanne.obj', b'/tmp/standin5.obj', 5, 20221018, 0.08, C.byref(n)); print('tris', n.value)
"
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=6 ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -14 gpurun_out/pytest_gpu.log
P="python scripts/profile_render.py"
for rep in 1 2; do $P --kernel bvh --spp 128 2>&1 | tail -1; done
$P --kernel bvh --spp 1024 2>&1 | tail -1
for st in 8 12 16 20 24; do RTW_LIB=raytracing-one-weekend_b200/librtw_b200_tune.so RTW_WF_STEPS=$st $P --kernel bvh --spp 128 2>&1 | tail -1 | sed "s/^/steps=$st /"; done
for rep in 1 2; do $P --kernel bvh --scene suzanne --spp 64 --depth 20 2>&1 | tail -1; done
for rep in 1 2; do $P --kernel bvh --scene /tmp/standin5.obj --spp 64 --depth 20 --bvh-build host 2>&1 | tail -1; done
