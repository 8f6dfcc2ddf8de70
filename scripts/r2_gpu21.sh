#!/bin/bash
# device BVH build: where to cut the radix tree (number of subtrees rebuilt with SAH on the device / size of the SAH top built on the host)
python -c "
import importlib,ctypes as C,sys
sys.path.insert(0,'.')
rtw=importlib.import_module('raytracing-one-weekend_b200')
n=C.c_longlong(0); rtw.host().rtwh_make_mesh(b'assets/suzanne.obj', b'/tmp/standin5.obj', 5, 20221018, 0.08, C.byref(n)); print('tris', n.value)
"
P="python scripts/profile_render.py --kernel bvh --scene /tmp/standin5.obj --spp 64 --depth 20 --bvh-build gpu"
$P > /dev/null 2>&1
for mc in 16384 4096 2048 512; do
  RTW_TRACE=1 RTW_LBVH_MAX_CLUSTERS=$mc $P 2>&1 | grep -E "device BVH|kernel" | grep -v launch | sed "s/^/max_clusters=$mc: /"
done
