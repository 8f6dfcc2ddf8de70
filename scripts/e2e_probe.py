import importlib, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
rtw = importlib.import_module("raytracing-one-weekend_b200")
aspect = 1.7777777777777777
scene = rtw.cover_scene(11, aspect)
for spp in (16, 16, 256, 1024):
    t0 = time.perf_counter()
    acc, st = rtw.render(scene, 1920, 1080, spp, 50)
    wall = (time.perf_counter() - t0) * 1e3
    print(f"spp {spp}: wall {wall:.1f} ms, total {st['total_ms']:.1f}, h2d/upload {st['h2d_ms']:.1f}, kernel {st['kernel_ms']:.1f}, d2h {st['d2h_ms']:.1f}")
