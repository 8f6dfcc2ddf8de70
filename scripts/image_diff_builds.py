#!/usr/bin/env python
"""Renders the cover frame with two builds of the CUDA library (subprocesses: RTW_LIB selects the build) and compares the accumulated
sums pixel by pixel: do two code generations of the same traversal produce the same picture?"""
import importlib, os, subprocess, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
if len(sys.argv) > 1 and sys.argv[1] == "--render":
    sys.path.insert(0, str(ROOT))
    rtw = importlib.import_module("raytracing-one-weekend_b200")
    if os.environ.get("RTW_LIB"):
        rtw.LIB_PATH = Path(os.environ["RTW_LIB"]).resolve()
    acc, st = rtw.render(rtw.cover_scene(11, 16 / 9), 1920, 1080, int(sys.argv[3]), 50)
    np.save(sys.argv[2], acc)
    print(sys.argv[2], st["paths"], st["rays"], st["kernel_ms"])
    sys.exit(0)
spp = sys.argv[1] if len(sys.argv) > 1 else "256"
outs = []
for tag, lib in (("product", ""), ("alt", sys.argv[2] if len(sys.argv) > 2 else "")):
    out = f"/tmp/img_{tag}.npy"
    subprocess.run([sys.executable, __file__, "--render", out, spp], env={**os.environ, "RTW_LIB": lib}, check=True)
    outs.append(np.load(out))
a, b = outs
d = np.abs(a.astype(np.float64) - b.astype(np.float64))
bad = d[..., :3].max(axis=2) > 0
print(f"pixels that differ: {int(bad.sum())} of {bad.size}; largest difference of a pixel sum {d.max():.6g} (sums are ~{a[..., :3].mean():.4g}); sample counts equal: {bool((a[..., 3] == b[..., 3]).all())}")
