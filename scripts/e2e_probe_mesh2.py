#!/usr/bin/env python
"""Host-side cost of a cold rtw_render call on the 991k-triangle stand-in, device build against host build, call by call."""
import ctypes as C, importlib, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
rtw = importlib.import_module("raytracing-one-weekend_b200")
n = C.c_longlong(0)
rtw.host().rtwh_make_mesh(str(Path(__file__).resolve().parents[1] / "assets/suzanne.obj").encode(), b"/tmp/standin5.obj", 5, 20221018, 0.08, C.byref(n))
sc = rtw.mesh_on_ground_scene("/tmp/standin5.obj", 16 / 9)
for label, fl in (("device", rtw.FLAG_BVH_BUILD_GPU), ("host", rtw.FLAG_BVH_BUILD_HOST), ("device", rtw.FLAG_BVH_BUILD_GPU)):
    for i in range(4):
        t = time.perf_counter()
        acc, st = rtw.render(sc, 1920, 1080, 16, 20, flags=rtw.FLAG_NO_SCENE_CACHE | fl)
        print(f"{label} call {i}: wall {1e3 * (time.perf_counter() - t):.1f} ms, flatten+build+upload {st['h2d_ms']:.1f}, device build {st['bvh_build_gpu_ms']:.1f}, kernel {st['kernel_ms']:.1f}, d2h {st['d2h_ms']:.1f}")
