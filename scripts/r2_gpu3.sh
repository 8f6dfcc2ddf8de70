#!/bin/bash
# A/B: binary 64-byte-node tree vs compressed 8-wide tree on the mesh workloads, no statistics, plus the CW kernel's knobs.
mkdir -p gpurun_out
python -c "
import importlib,ctypes as C,sys
sys.path.insert(0,'.')
rtw=importlib.import_module('raytracing-one-weekend_b200')
n=C.c_longlong(0); rtw.host().rtwh_make_mesh(b'assets/suzanne.obj', b'/tmp/standin5.obj', 5, 20221018, 0.08, C.byref(n)); print('tris', n.value)
"
R="python scripts/profile_render.py --kernel bvh --scene /tmp/standin5.obj --spp 64 --depth 20"
RTW_MESH_BVH=binary $R 2>&1 | tail -1 | sed "s/^/binary /"
RTW_MESH_BVH=binary $R 2>&1 | tail -1 | sed "s/^/binary /"
for minb in 3 4 2; do for steps in 1 2 4; do for svc in 8 16 20 24; do
  RTW_CW_LEAF=1 RTW_LEAF_MIN=1 RTW_CW_MINB=$minb RTW_CW_STEPS=$steps RTW_CW_SERVICE=$svc $R 2>&1 | tail -1 | cut -d: -f2 | cut -d, -f1-2 | sed "s/^/cw minb=$minb steps=$steps svc=$svc /"
done; done; done
S="python scripts/profile_render.py --kernel bvh --scene suzanne --spp 64 --depth 20"
RTW_MESH_BVH=binary $S 2>&1 | tail -1 | sed "s/^/binary /"
for steps in 1 2 4; do for svc in 8 16 20 24; do
  RTW_CW_LEAF=1 RTW_LEAF_MIN=1 RTW_CW_STEPS=$steps RTW_CW_SERVICE=$svc $S 2>&1 | tail -1 | cut -d: -f2 | cut -d, -f1-2 | sed "s/^/cw suzanne steps=$steps svc=$svc /"
done; done
# ncu: the CW kernel on the stand-in (compare with profiles/r01_prof_k2_dragon.txt, the binary tree on the same frame)
P="--kernel bvh --scene /tmp/standin5.obj --spp 4 --depth 20"
RTW_CW_LEAF=1 RTW_LEAF_MIN=1 python scripts/profile_render.py $P > gpurun_out/plain_cw_dragon.log 2>&1 &&
RTW_CW_LEAF=1 RTW_LEAF_MIN=1 ncu --set full --clock-control none --import-source on -k regex:k_render -s 1 -c 1 -f -o gpurun_out/prof_cw_dragon python scripts/profile_render.py $P > gpurun_out/ncu_cw_dragon.log 2>&1
RTW_MESH_BVH=binary python scripts/profile_render.py $P > gpurun_out/plain_k2_dragon.log 2>&1 &&
RTW_MESH_BVH=binary ncu --set full --clock-control none --import-source on -k regex:k_render -s 1 -c 1 -f -o gpurun_out/prof_k2_dragon python scripts/profile_render.py $P > gpurun_out/ncu_k2_dragon.log 2>&1
ls -la gpurun_out/*.ncu-rep
