#!/bin/bash
# Round 2, second GPU call: the compressed wide BVH on the mesh workloads (parity subset, then speed + knobs).
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x -k "mesh or suzanne or stand_in or 991k or edge or mixed or primary" ) > gpurun_out/pytest_mesh.log 2>&1; tail -12 gpurun_out/pytest_mesh.log
python -c "
import importlib,ctypes as C,sys
sys.path.insert(0,'.')
rtw=importlib.import_module('raytracing-one-weekend_b200')
n=C.c_longlong(0); rtw.host().rtwh_make_mesh(b'assets/suzanne.obj', b'/tmp/standin5.obj', 5, 20221018, 0.08, C.byref(n)); print('tris', n.value)
"
for leaf in 1 2 3; do for lm in 1 2 4 8; do
  RTW_CW_LEAF=$leaf RTW_LEAF_MIN=$lm python scripts/profile_render.py --kernel bvh --scene /tmp/standin5.obj --spp 64 --depth 20 --stats 2>&1 | tail -1 | sed "s/^/leaf=$leaf leaf_min=$lm /"
done; done
for leaf in 1 2 3; do for lm in 1 4; do
  RTW_CW_LEAF=$leaf RTW_LEAF_MIN=$lm python scripts/profile_render.py --kernel bvh --scene suzanne --spp 64 --depth 20 --stats 2>&1 | tail -1 | sed "s/^/leaf=$leaf leaf_min=$lm /"
done; done
