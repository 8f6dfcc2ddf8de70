#!/bin/bash
# same-box A/B: packed FFMA2 slab test (product build) against the scalar form (librtw_b200_scalar.so), and the warp-count sensitivity of K2w
mkdir -p gpurun_out
python -c "
import importlib,ctypes as C,sys
sys.path.insert(0,'.')
rtw=importlib.import_module('raytracing-one-weekend_b200')
n=C.c_longlong(0); rtw.host().rtwh_make_mesh(b'assets/suzanne.obj', b'/tmp/standin5.obj', 5, 20221018, 0.08, C.byref(n)); print('tris', n.value)
"
P="python scripts/profile_render.py"
for rep in 1 2; do
for lib in "" raytracing-one-weekend_b200/librtw_b200_scalar.so; do
  tag=${lib:+scalar}; tag=${tag:-ffma2}
  RTW_LIB=$lib $P --kernel bvh --spp 128 2>&1 | tail -1 | sed "s/^/$tag /"
  RTW_LIB=$lib $P --kernel bvh-perlane --spp 128 2>&1 | tail -1 | sed "s/^/$tag /"
  RTW_LIB=$lib $P --kernel bvh --scene suzanne --spp 64 --depth 20 2>&1 | tail -1 | sed "s/^/$tag /"
  RTW_LIB=$lib $P --kernel bvh --scene /tmp/standin5.obj --spp 64 --depth 20 --bvh-build host 2>&1 | tail -1 | sed "s/^/$tag host-built /"
done; done
for w in 20 24 28 32; do for rep in 1 2; do RTW_WF_WARPS=$w $P --kernel bvh --spp 128 2>&1 | tail -1 | sed "s/^/warps=$w /"; done; done
RTW_WF_WARPS=32 $P --kernel bvh --spp 1024 2>&1 | tail -1 | sed "s/^/warps=32 /"
RTW_WF_WARPS=32 timeout 600 python -m pytest tests -m gpu -q -x -k "cover or determinism or depth_rule or full_size" 2>&1 | tail -3
