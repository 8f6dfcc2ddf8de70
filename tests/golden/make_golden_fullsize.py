#!/usr/bin/env python
"""Full-size golden for BASELINE config 2: the UNMODIFIED reference (oracle/_ref) renders the cover scene at
1920x1080 x 1024 spp x depth 50 (`-w 1920 -a 1.7777777777777777 -s 1024 -c 50`, motion blur on), the sentence the
north star ends on.  Build container only (/root/reference is needed to build oracle/_ref); ~25 min on 8 cores.

    nice python tests/golden/make_golden_fullsize.py [--procs 8] [--chunk 16]

The 1024 samples are rendered as 64 single-threaded chunks of 16 spp (render.cpp:152-163 looping over ray_color on the
global mt19937, reseeded per chunk with 1000 + 7919*i so that chunks are independent: SURVEY Q9) and summed, which is
what render.cpp:169-180 does with its threads, minus the data race.  Stored in cover_1080p_1024spp_depth50.npz:

  rgb        uint8 [1080,1920,3]   write_color (render.cpp:11-20) of the summed image: what the reference would print
  mean_ch    float64 [3]           image mean per channel (linear domain)
  se_ch      float64 [3]           standard error of that mean (from the per-pixel sample variances)
  blk_mean   float32 [135,240,3]   linear mean over 8x8 pixel blocks
  blk_se     float32 [135,240,3]   standard error of the block means
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
OUT = Path(__file__).resolve().parent

W, H, SPP, DEPTH, ASPECT = 1920, 1080, 1024, 50, 1.7777777777777777


def _worker(args):
    import oracle
    i, chunk = args
    ref = oracle.ref()
    sc = ref.scene_cover(11, ASPECT, True, seed=5489)
    s, q, _ = sc.render_linear(W, H, chunk, DEPTH, seed=1000 + 7919 * i)
    return i, s, q


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=8)
    ap.add_argument("--chunk", type=int, default=16)
    a = ap.parse_args()
    import oracle
    oracle.build()
    nchunks = SPP // a.chunk
    s = np.zeros((H, W, 3))
    q = np.zeros((H, W, 3))
    t0 = time.time()
    with mp.Pool(a.procs) as pool:
        for k, (i, si, qi) in enumerate(pool.imap_unordered(_worker, [(i, a.chunk) for i in range(nchunks)])):
            s += si
            q += qi
            print(f"chunk {i} done ({k + 1}/{nchunks}) {time.time() - t0:.0f}s", flush=True)
    n = nchunks * a.chunk
    mean = s / n
    var = np.maximum(q / n - mean * mean, 0.0) * n / (n - 1)       # per-sample variance per pixel and channel
    port = oracle.port()
    rgb = port.quantize(s, n)                                       # write_color, pinned against the reference in test_oracle_pin
    mean_ch = mean.mean(axis=(0, 1))
    se_ch = np.sqrt(var.sum(axis=(0, 1)) / n) / (W * H)
    bm = mean.reshape(H // 8, 8, W // 8, 8, 3).mean(axis=(1, 3))
    bse = np.sqrt(var.reshape(H // 8, 8, W // 8, 8, 3).sum(axis=(1, 3)) / n) / 64.0
    meta = {"generator": "tests/golden/make_golden_fullsize.py",
            "source": "oracle/_ref (reference sources unmodified, glm/CLI11/tinyobj/fmt shims)",
            "scene": "cover (lots_of_balls, nsqrt 11, moving spheres)", "width": W, "height": H, "spp": n, "max_child_rays": DEPTH,
            "aspect": ASPECT, "how": f"{nchunks} single-threaded runs of render.cpp:152-163 of {a.chunk} spp, seeds 1000+7919*i, summed",
            "cpu_seconds_wall": round(time.time() - t0)}
    np.savez_compressed(OUT / "cover_1080p_1024spp_depth50.npz", rgb=rgb, mean_ch=mean_ch, se_ch=se_ch,
                        blk_mean=bm.astype(np.float32), blk_se=bse.astype(np.float32), meta=json.dumps(meta))
    print("mean", mean_ch, "se", se_ch, "wall", time.time() - t0)


if __name__ == "__main__":
    main()
