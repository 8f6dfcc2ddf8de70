#!/usr/bin/env python
"""Which render kernel is fastest at which scene size (cover scene with -n N, 1080p, 64 spp, one GPU): the shared-memory sphere
sweep (K1), the BVH kernels as the library picks them (K2w while tables + records fit in shared memory, K2 beyond), and the per-lane
BVH kernel forced (K2).    python scripts/kernel_crossover.py [N ...]"""
import importlib
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
rtw = importlib.import_module("raytracing-one-weekend_b200")
aspect = 1.7777777777777777
for n in [int(x) for x in sys.argv[1:]] or [1, 2, 3, 4, 6, 11, 12, 13, 14, 15, 16, 17, 20]:
    sc = rtw.cover_scene(n, aspect)
    out = []
    for name, k in (("K1 spheres", rtw.KERNEL_SPHERES_SMEM), ("BVH auto", rtw.KERNEL_BVH), ("BVH per-lane", rtw.KERNEL_BVH_PERLANE)):
        best = None
        try:
            for i in range(3):
                acc, st = rtw.render(sc, 1920, 1080, 64, 50, kernel=k)
                best = st["kernel_ms"] if best is None else min(best, st["kernel_ms"])
            out.append(f"{name} {st['paths'] / best / 1e3:8.1f}" + (f" (variant {st['bvh_variant']})" if name == "BVH auto" else ""))
        except rtw.RtwError:
            out.append(f"{name}      n/a")
    print(f"-n {n}: {len(sc.prims):5d} primitives  " + "   ".join(out), flush=True)
