#!/usr/bin/env python
"""Host-side cost of rtw_scene_upload on the 991,232-triangle stand-in: flatten + SAH build (rtw_flatten_info, no GPU call)."""
import ctypes as C
import importlib
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
rtw = importlib.import_module("raytracing-one-weekend_b200")
n = C.c_longlong(0)
rtw.host().rtwh_make_mesh(str(Path(__file__).resolve().parents[1] / "assets/suzanne.obj").encode(), b"/tmp/standin5.obj", 5, 20221018, 0.08, C.byref(n))
t = time.time(); sc = rtw.mesh_on_ground_scene("/tmp/standin5.obj", 1.7777777777777777); print(f"OBJ parse + scene: {time.time() - t:.3f} s, {os.cpu_count()} cpus")
for i in range(4):
    t = time.time(); r = rtw.flatten_info(sc)
    print(f"flatten_info {1e3 * (time.time() - t):.1f} ms: flatten {r['flatten_ms']:.1f} ms of which BVH build {r['bvh_build_ms']:.1f} ms, depth {r['bvh_max_depth']}, nodes {r['n_bvh_nodes']}")
