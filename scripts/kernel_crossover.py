import importlib, sys
sys.path.insert(0, '.')
rtw = importlib.import_module("raytracing-one-weekend_b200")
aspect = 1.7777777777777777
for n in (1, 2, 3, 4, 5, 6):
    sc = rtw.cover_scene(n, aspect)
    out = []
    for name, k in (("spheres", rtw.KERNEL_SPHERES_SMEM), ("bvh", rtw.KERNEL_BVH), ("auto", rtw.KERNEL_AUTO)):
        best = None
        for i in range(3):
            acc, st = rtw.render(sc, 1920, 1080, 64, 50, kernel=k)
            best = st["kernel_ms"] if best is None else min(best, st["kernel_ms"])
        out.append(f"{name} {st['paths'] / best / 1e3:8.1f}")
    print(f"nsqrt {n}: {len(sc.prims):4d} prims  " + "  ".join(out) + f"   (auto used kernel {st['kernel_used']} variant {st['bvh_variant']})", flush=True)
