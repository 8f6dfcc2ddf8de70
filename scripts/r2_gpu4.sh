#!/bin/bash
# mesh kernel (binary tree): tree-ordered triangle records, L2 residency window, evict-first REDs -- A/B on the stand-in and suzanne
mkdir -p gpurun_out
python -c "
import importlib,ctypes as C,sys
sys.path.insert(0,'.')
rtw=importlib.import_module('raytracing-one-weekend_b200')
n=C.c_longlong(0); rtw.host().rtwh_make_mesh(b'assets/suzanne.obj', b'/tmp/standin5.obj', 5, 20221018, 0.08, C.byref(n)); print('tris', n.value)
"
R="python scripts/profile_render.py --kernel bvh --scene /tmp/standin5.obj --spp 64 --depth 20"
for t in insertion leaf; do for l2 in 0 1; do for red in 0 1; do
  RTW_TRI_ORDER=$t RTW_L2_PERSIST=$l2 RTW_RED_HINT=$red $R 2>&1 | tail -1 | cut -d: -f2 | cut -d, -f1-2 | sed "s/^/standin tri=$t l2persist=$l2 redhint=$red /"
done; done; done
RTW_L2_PERSIST=0 RTW_RED_HINT=0 $R --width 960 2>&1 | tail -1 | sed "s/^/960px nohint /"
$R --width 960 2>&1 | tail -1 | sed "s/^/960px hints /"
S="python scripts/profile_render.py --kernel bvh --scene suzanne --spp 64 --depth 20"
for t in insertion leaf; do RTW_TRI_ORDER=$t $S 2>&1 | tail -1 | cut -d: -f2 | cut -d, -f1-2 | sed "s/^/suzanne tri=$t /"; done
# host build time on this box: flatten + SAH build of the 991k-triangle mesh (both tree formats), 3 repeats
python - <<'PY'
import importlib, time, os, sys
sys.path.insert(0,'.')
r=importlib.import_module('raytracing-one-weekend_b200')
sc=r.mesh_on_ground_scene('/tmp/standin5.obj',16/9)
for env in ({}, {"RTW_MESH_BVH":"cw8"}, {"RTW_SAH_SINGLE_AXIS_BELOW":"4096"}):
    for k in ("RTW_MESH_BVH","RTW_SAH_SINGLE_AXIS_BELOW"): os.environ.pop(k,None)
    os.environ.update(env)
    for i in range(3):
        t=time.time(); inf=r.flatten_info(sc); dt=time.time()-t
    print(env, "flatten_ms", round(inf['flatten_ms'],1), "bvh_ms", round(inf['bvh_build_ms'],1), "cores", os.cpu_count())
t=time.time(); h=r.scene_hash(sc); print("hash ms", round((time.time()-t)*1e3,2))
PY
timeout 900 python -m pytest tests -m gpu -q -x -k "mesh or suzanne or stand_in or 991k or mixed" 2>&1 | tail -4
