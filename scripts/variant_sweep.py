#!/usr/bin/env python
"""Kernel-variant sweep on one GPU: every variant renders the same frame; prints throughput and the difference of its accumulation
buffer from the first variant's.  Variant ids: 0 = BVH (library's choice), 100 = BVH per-lane kernel forced, 1 = sphere sweep;
any other id is passed through rtw_render_cfg.rays_per_lane to whatever experimental instantiations the library was built with.
    python scripts/variant_sweep.py --variants 0,100 [--scene cover|suzanne|standin|PATH.obj] [--spp 64]"""
import argparse
import importlib
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
rtw = importlib.import_module("raytracing-one-weekend_b200")

ap = argparse.ArgumentParser()
ap.add_argument("--variants", default="0")
ap.add_argument("--spp", type=int, default=64)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--depth", type=int, default=50)
ap.add_argument("--scene", default="cover")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
aspect = 1.7777777777777777
if a.scene == "cover":
    scene = rtw.cover_scene(11, aspect)
elif a.scene == "suzanne":
    scene = rtw.mesh_on_ground_scene(str(Path(__file__).resolve().parents[1] / "assets/suzanne.obj"), aspect)
elif a.scene == "standin":  # 991,232-triangle stand-in for dragon.obj (5 subdivision rounds of suzanne)
    import ctypes as C
    n = C.c_longlong(0)
    if rtw.host().rtwh_make_mesh(str(Path(__file__).resolve().parents[1] / "assets/suzanne.obj").encode(), b"/tmp/standin5.obj", 5, 20221018, 0.08, C.byref(n)) != 0:
        raise RuntimeError(rtw.host().rtwh_last_error().decode())
    scene = rtw.mesh_on_ground_scene("/tmp/standin5.obj", aspect)
else:
    scene = rtw.mesh_on_ground_scene(a.scene, aspect)
H = rtw.image_height(a.width, aspect)
base = None
for v in [int(x) for x in a.variants.split(",")]:
    best = None
    try:
        for i in range(a.reps):
            kw = {100: dict(kernel=rtw.KERNEL_BVH_PERLANE), 1: dict(kernel=rtw.KERNEL_SPHERES_SMEM)}.get(v, dict(kernel=rtw.KERNEL_BVH, rays_per_lane=v))
            acc, st = rtw.render(scene, a.width, H, a.spp, a.depth, **kw)
            best = st["kernel_ms"] if best is None else min(best, st["kernel_ms"])
    except Exception as e:  # noqa: BLE001
        print(f"variant {v}: FAILED {e}")
        continue
    if base is None:
        base = acc
    diff = np.abs(acc[..., :3] - base[..., :3])
    p, r = st["paths"], st["rays"]
    print(f"{a.scene} variant {v:4d} {a.width}x{H}x{a.spp}: {best:8.2f} ms {p / best / 1e3:8.1f} Mpaths/s {r / best / 1e3:8.1f} Mrays/s rays/path {r / p:.4f} "
          f"paths_ok {bool(np.all(acc[..., 3] == a.spp))} max|diff| {diff.max() / a.spp:.3e} mean|diff| {diff.mean() / a.spp:.3e} pixels_diff {(diff.max(axis=2) > 0).mean():.4f}", flush=True)
