#!/usr/bin/env python
"""Tiny driver for ncu: upload the cover scene, render twice (the second launch is the one to profile).
    python scripts/profile_render.py --kernel spheres --rays-per-lane 4 --spp 16 [--width 1920]"""
import argparse
import importlib
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
rtw = importlib.import_module("raytracing-one-weekend_b200")
import os
if os.environ.get("RTW_LIB"):   # A/B builds of the CUDA library (csrc/build.sh with RTW_OUT=...): a tuning aid of this script only
    rtw.LIB_PATH = Path(os.environ["RTW_LIB"]).resolve()

ap = argparse.ArgumentParser()
ap.add_argument("--kernel", default="auto", choices=["auto", "spheres", "bvh", "bvh-perlane"])
ap.add_argument("--rays-per-lane", type=int, default=0, help="sphere sweep only: 1, 2 or 4 paths per lane")
ap.add_argument("--spp", type=int, default=16)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--depth", type=int, default=50)
ap.add_argument("--scene", default="cover")
ap.add_argument("--stats", action="store_true")
ap.add_argument("--bvh-build", default="auto", choices=["auto", "host", "gpu"])
a = ap.parse_args()
aspect = 1.7777777777777777
if a.scene == "cover":
    scene = rtw.cover_scene(11, aspect)
elif a.scene == "suzanne":
    scene = rtw.mesh_on_ground_scene(str(Path(__file__).resolve().parents[1] / "assets/suzanne.obj"), aspect)
else:
    scene = rtw.mesh_on_ground_scene(a.scene, aspect)
H = rtw.image_height(a.width, aspect)
k = {"auto": rtw.KERNEL_AUTO, "spheres": rtw.KERNEL_SPHERES_SMEM, "bvh": rtw.KERNEL_BVH, "bvh-perlane": rtw.KERNEL_BVH_PERLANE}[a.kernel]
for i in range(2):
    acc, st = rtw.render(scene, a.width, H, a.spp, a.depth, kernel=k, rays_per_lane=a.rays_per_lane, stats=a.stats,
                         flags={"auto": 0, "host": rtw.FLAG_BVH_BUILD_HOST, "gpu": rtw.FLAG_BVH_BUILD_GPU}[a.bvh_build])
p, r = st["paths"], st["rays"]
print(f"# launch: scene={a.scene if '/' not in a.scene else 'standin'} kernel={a.kernel} width={a.width} height={H} spp={a.spp} depth={a.depth} paths={p} rays={r}")
print(f"{a.scene} {a.kernel} rpl={a.rays_per_lane} {a.width}x{H}x{a.spp}: kernel {st['kernel_ms']:.2f} ms, {p / st['kernel_ms'] / 1e3:.1f} Mpaths/s, "
      f"{r / st['kernel_ms'] / 1e3:.1f} Mrays/s, rays/path {r / p:.3f}"
      + (f", per ray: sphere tests {st['sphere_tests'] / r:.2f} candidates {st['sphere_candidates'] / r:.2f} nodes {st['node_visits'] / r:.2f} tris {st['tri_tests'] / r:.2f}" if a.stats else ""))
