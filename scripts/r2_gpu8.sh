#!/bin/bash
# FFMA2 (packed fp32 FMA) in the node slab test, nodes with (left, right) pairs: probe of the instruction, parity, speed on cover / suzanne / stand-in
mkdir -p gpurun_out
./scripts/ffma2_probe 2>&1 | tee gpurun_out/ffma2_probe.txt
python -c "
import importlib,ctypes as C,sys
sys.path.insert(0,'.')
rtw=importlib.import_module('raytracing-one-weekend_b200')
n=C.c_longlong(0); rtw.host().rtwh_make_mesh(b'assets/suzanne.obj', b'/tmp/standin5.obj', 5, 20221018, 0.08, C.byref(n)); print('tris', n.value)
"
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for i in 1 2; do python scripts/profile_render.py --kernel bvh --spp 128 2>&1 | tail -1; done
python scripts/profile_render.py --kernel bvh --spp 1024 2>&1 | tail -1
python scripts/profile_render.py --kernel bvh-perlane --spp 128 2>&1 | tail -1
for i in 1 2; do python scripts/profile_render.py --kernel bvh --scene suzanne --spp 64 --depth 20 2>&1 | tail -1; done
for i in 1 2; do python scripts/profile_render.py --kernel bvh --scene /tmp/standin5.obj --spp 64 --depth 20 2>&1 | tail -1; done
