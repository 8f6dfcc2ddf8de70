#!/bin/bash
# K2w with 4 visits per leaf check x 5 per phase (product); per-lane K2 with 2 visits per leaf phase (A/B build)
python -c "
import importlib,ctypes as C,sys
sys.path.insert(0,'.')
rtw=importlib.import_module('raytracing-one-weekend_b200')
n=C.c_longlong(0); rtw.host().rtwh_make_mesh(b'assets/suzanne.obj', b'/tmp/standin5.obj', 5, 20221018, 0.08, C.byref(n)); print('tris', n.value)
"
P="python scripts/profile_render.py"
$P --kernel bvh --spp 128 2>&1 | tail -1
$P --kernel bvh --spp 1024 2>&1 | tail -1
for rep in 1 2; do
for tag in product k2v2; do
  lib=raytracing-one-weekend_b200/librtw_b200_$tag.so; [ $tag = product ] && lib=""
  RTW_LIB=$lib $P --kernel bvh-perlane --spp 128 2>&1 | tail -1 | sed "s/^/$tag /"
  RTW_LIB=$lib $P --kernel bvh --scene suzanne --spp 64 --depth 20 2>&1 | tail -1 | sed "s/^/$tag /"
  RTW_LIB=$lib $P --kernel bvh --scene /tmp/standin5.obj --spp 64 --depth 20 --bvh-build host 2>&1 | tail -1 | sed "s/^/$tag /"
done; done
timeout 900 python -m pytest tests -m gpu -q -x -k "cover or determinism or depth_rule or full_size or crossover or grids" 2>&1 | tail -3
