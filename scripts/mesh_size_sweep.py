#!/usr/bin/env python
"""Per-lane BVH kernel against an experimental variant id on suzanne subdivided 0..R times (968 * 4^k triangles on the ground sphere).
    python scripts/mesh_size_sweep.py VARIANT [R=4]"""
import ctypes as C
import importlib
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
rtw = importlib.import_module("raytracing-one-weekend_b200")
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
R = int(sys.argv[2]) if len(sys.argv) > 2 else 4
aspect = 1.7777777777777777
for k in range(R + 1):
    path = str(ROOT / "assets/suzanne.obj")
    if k > 0:
        out = f"/tmp/suz_r{k}.obj"
        n = C.c_longlong(0)
        rtw.host().rtwh_make_mesh(path.encode(), out.encode(), k, 20221018, 0.08, C.byref(n))
        path = out
    sc = rtw.mesh_on_ground_scene(path, aspect)
    res = []
    for v in (0, variant):
        best = None
        for i in range(3):
            acc, st = rtw.render(sc, 1920, 1080, 32, 20, kernel=rtw.KERNEL_BVH, rays_per_lane=v)
            best = st["kernel_ms"] if best is None else min(best, st["kernel_ms"])
        res.append(st["paths"] / best / 1e3)
    print(f"rounds {k}: {len(sc.prims) - 1:8d} triangles   default {res[0]:8.1f}   variant {variant} {res[1]:8.1f} Mpaths/s ({100 * (res[1] / res[0] - 1):+.1f} %)", flush=True)
