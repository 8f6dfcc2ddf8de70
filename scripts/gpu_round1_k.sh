#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/stats.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
for v in 0 101 102 103 104; do
  python scripts/profile_render.py --kernel bvh --rays-per-lane $v --spp 128 >> gpurun_out/stats.log 2>&1
done
python scripts/profile_render.py --kernel bvh --spp 32 --stats >> gpurun_out/stats.log 2>&1
python scripts/profile_render.py --kernel bvh --scene suzanne --spp 128 >> gpurun_out/stats.log 2>&1
RTW_BVH_LEAF=2 python scripts/profile_render.py --kernel bvh --scene suzanne --spp 128 >> gpurun_out/stats.log 2>&1
python -c "
import importlib,ctypes as C,sys
sys.path.insert(0,'.')
rtw=importlib.import_module('raytracing-one-weekend_b200')
n=C.c_longlong(0); rtw.host().rtwh_make_mesh(b'tests/golden/suzanne.obj', b'/tmp/standin5.obj', 5, 20221018, 0.08, C.byref(n)); print('tris', n.value)
"
for leaf in 1 2 4; do
  echo "dragon leaf=$leaf" >> gpurun_out/stats.log
  RTW_BVH_LEAF=$leaf python scripts/profile_render.py --kernel bvh --scene /tmp/standin5.obj --spp 64 --depth 20 --stats >> gpurun_out/stats.log 2>&1
  RTW_BVH_LEAF=$leaf python scripts/profile_render.py --kernel bvh --scene /tmp/standin5.obj --spp 64 --depth 20 >> gpurun_out/stats.log 2>&1
done
grep -v "^Scene has" gpurun_out/stats.log
