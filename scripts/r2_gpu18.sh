#!/bin/bash
# materials next to the spheres (one L2 round trip less per shading batch), ground sphere in the kernel parameters: parity + speed
mkdir -p gpurun_out
python -c "
import importlib,ctypes as C,sys
sys.path.insert(0,'.')
rtw=importlib.import_module('raytracing-one-weekend_b200')
n=C.c_longlong(0); rtw.host().rtwh_make_mesh(b'assets/suzanne.obj', b'/tmp/standin5.obj', 5, 20221018, 0.08, C.byref(n)); print('tris', n.value)
"
( timeout 1500 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -3
P="python scripts/profile_render.py"
for rep in 1 2; do $P --kernel bvh --spp 128 2>&1 | tail -1; done
$P --kernel bvh --spp 1024 2>&1 | tail -1
$P --kernel bvh-perlane --spp 128 2>&1 | tail -1
$P --kernel spheres --spp 64 2>&1 | tail -1
for rep in 1 2; do $P --kernel bvh --scene suzanne --spp 64 --depth 20 2>&1 | tail -1; done
for rep in 1 2; do $P --kernel bvh --scene /tmp/standin5.obj --spp 64 --depth 20 --bvh-build host 2>&1 | tail -1; done
