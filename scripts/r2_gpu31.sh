#!/bin/bash
# per-lane K2 with two visits per leaf phase: lanes that must stand on a leaf before triangle leaves are tested (RTW_LEAF_MIN), meshes
python -c "
import importlib,ctypes as C,sys
sys.path.insert(0,'.')
rtw=importlib.import_module('raytracing-one-weekend_b200')
n=C.c_longlong(0); rtw.host().rtwh_make_mesh(b'assets/suzanne.obj', b'/tmp/standin5.obj', 5, 20221018, 0.08, C.byref(n)); print('tris', n.value)
"
P="python scripts/profile_render.py"
for lm in 4 1 2 3 6 8; do
  RTW_LEAF_MIN=$lm $P --kernel bvh --scene suzanne --spp 64 --depth 20 2>&1 | tail -1 | sed "s/^/leaf_min=$lm /"
  RTW_LEAF_MIN=$lm $P --kernel bvh --scene /tmp/standin5.obj --spp 64 --depth 20 --bvh-build host 2>&1 | tail -1 | sed "s/^/leaf_min=$lm /"
done
