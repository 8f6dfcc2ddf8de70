#!/bin/bash
# device BVH build: SAH inside the clusters (one warp per subtree) against the radix subtrees -- tree check, trace speed, build time
mkdir -p gpurun_out
python -c "
import importlib,ctypes as C,sys
sys.path.insert(0,'.')
rtw=importlib.import_module('raytracing-one-weekend_b200')
n=C.c_longlong(0); rtw.host().rtwh_make_mesh(b'assets/suzanne.obj', b'/tmp/standin5.obj', 5, 20221018, 0.08, C.byref(n)); print('tris', n.value)
"
timeout 900 python -m pytest tests -m gpu -q -x -k "built_on_the_device or 991k or stand_in" 2>&1 | tail -3
P="python scripts/profile_render.py --kernel bvh --scene /tmp/standin5.obj --spp 64 --depth 20"
for rep in 1 2; do
RTW_TRACE=1 RTW_LBVH_SAH_CLUSTERS=0 $P --bvh-build gpu 2>&1 | grep -E "device BVH|kernel" | sed "s/^/radix clusters: /"
RTW_TRACE=1 $P --bvh-build gpu 2>&1 | grep -E "device BVH|kernel" | sed "s/^/SAH clusters:   /"
$P --bvh-build host 2>&1 | tail -1 | sed "s/^/host SAH:       /"
done
