#!/usr/bin/env python
"""Generates the golden fixtures in tests/golden/ from the UNMODIFIED reference (oracle/_ref, built by
`make -C oracle ref` from /root/reference/src).  Run in the build container only: /root/reference does not exist on
the GPU box.  Every fixture records how it was made in its `meta` entry.

    python tests/golden/make_golden.py            # everything (a few minutes on 8 cores)
"""
from __future__ import annotations

import hashlib
import json
import multiprocessing as mp
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
OUT = Path(__file__).resolve().parent
SUZANNE = "/root/reference/suzanne.obj"


def _linear_worker(args):
    """One reference process-equivalent: seed the global mt19937, render `spp` samples on one thread."""
    import oracle
    kind, width, height, spp, depth, seed, aspect = args
    ref = oracle.ref()
    # the scene is built from the default-seeded stream (as a fresh reference process would), THEN the stream is reseeded
    if kind == "suzanne_on_ground":
        # BASELINE config 3 as SURVEY 8(d) defines it (3b): the reference has no such scene builder (Q14), so the product's own
        # builder lays out the primitives and the REFERENCE renders them through its public Scene API (ref_scene_custom)
        import importlib
        rtw = importlib.import_module("raytracing-one-weekend_b200")
        ms = rtw.mesh_on_ground_scene(SUZANNE, aspect)
        sc = ref.scene_custom(ms.prims, ms.mats.view(oracle.MAT_DTYPE), oracle.camera_params(**ms.params))
    else:
        sc = ref.scene_cover(11, aspect, True, seed=5489) if kind == "cover" else \
            ref.scene_cover(11, aspect, False, seed=5489) if kind == "cover_static" else ref.scene_obj(SUZANNE, aspect, seed=5489)
    s, q, _ = sc.render_linear(width, height, spp, depth, seed=seed)
    return s, q


def converged(kind, width, height, spp_total, depth, aspect, nproc=8):
    per = spp_total // nproc
    jobs = [(kind, width, height, per, depth, 1000 + 7919 * i, aspect) for i in range(nproc)]
    with mp.Pool(nproc) as pool:
        res = pool.map(_linear_worker, jobs)
    s = sum(r[0] for r in res)
    q = sum(r[1] for r in res)
    n = per * nproc
    mean = s / n
    var = np.maximum(q / n - mean * mean, 0.0) * n / (n - 1)
    return mean.astype(np.float32), var.astype(np.float32), n


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="", help="regenerate only the converged fixture of this name")
    only = ap.parse_args().only
    import oracle
    oracle.build()
    ref = oracle.ref()
    meta_common = {"generator": "tests/golden/make_golden.py", "source": "oracle/_ref (reference sources unmodified, glm/CLI11/tinyobj/fmt shims)"}

    if not only:
        primary_and_default(ref, oracle, meta_common)
    converged_fixtures(meta_common, only)


def primary_and_default(ref, oracle, meta_common):
    # 1. the reference executable's own output, default configuration, one thread (bit-reproducible: SURVEY Q9)
    ppm = subprocess.run([str(oracle.REF_EXE), "-t", "1"], capture_output=True, check=True).stdout
    md5 = hashlib.md5(ppm).hexdigest()
    tok = ppm.split()
    w, h = int(tok[1]), int(tok[2])
    img = np.array(tok[4:], dtype=np.int64).astype(np.uint8).reshape(h, w, 3)
    np.savez_compressed(OUT / "cover_default_t1.npz", rgb=img,
                        meta=json.dumps({**meta_common, "cmd": "rtweekend_ref -t 1", "md5_of_p3_text": md5, "width": w, "height": h,
                                         "spp": 20, "max_child_rays": 20}))
    print("cover_default_t1", md5)

    # 2. primary hits: cover scene (time 0 and 0.5) and suzanne `foo` scene
    sc = ref.scene_cover(11, 1.5, True)
    for time in (0.0, 0.5):
        pid, t, nrm, front = sc.primary_hits(200, 133, time)
        assert sc.bvh_vs_bruteforce_disagreements == 0
        np.savez_compressed(OUT / f"cover_primary_200x133_t{time:g}.npz", id=pid.astype(np.int16), t=t, normal=nrm.astype(np.float32),
                            front=front, meta=json.dumps({**meta_common, "scene": "lots_of_balls default", "time": time,
                                                          "mode": "aperture 0, shutter [time,time], pixel centres"}))
    so = ref.scene_obj(SUZANNE, 1.5)
    pid, t, nrm, front = so.primary_hits(200, 133, 0.0)
    assert so.bvh_vs_bruteforce_disagreements == 0
    np.savez_compressed(OUT / "suzanne_primary_200x133.npz", id=pid.astype(np.int16), t=t, normal=nrm.astype(np.float32), front=front,
                        meta=json.dumps({**meta_common, "scene": "foo(suzanne.obj)", "time": 0.0}))
    print("primary hits done")


def converged_fixtures(meta_common, only):
    # 3. converged linear-domain statistics (mean and per-sample variance per pixel and channel)
    converged_sets = [("cover_converged_120x80", "cover", 120, 80, 1024, 20, 1.5),
                      ("cover_static_converged_96x54", "cover_static", 96, 54, 512, 50, 1.7777777777777777),
                      ("suzanne_converged_96x64", "suzanne", 96, 64, 512, 20, 1.5),
                      # BASELINE config 2's scene, camera and depth (cover, 16:9, depth 50, motion blur) at a fifth of its resolution
                      # and its full 1024 spp: the largest reference render that stays a small fixture (variance stored as float16)
                      ("cover_converged_384x216_depth50", "cover", 384, 216, 1024, 50, 1.7777777777777777),
                      # BASELINE config 3 (3b of SURVEY 8(d)): suzanne on the r=1000 ground, 16:9, depth 20, at a sixth of its resolution
                      ("suzanne_on_ground_converged_320x180", "suzanne_on_ground", 320, 180, 1024, 20, 1.7777777777777777)]
    for name, kind, w, h, spp, depth, aspect in converged_sets:
        if only and name != only:
            continue
        mean, var, n = converged(kind, w, h, spp, depth, aspect)
        if w * h > 20000:
            var = var.astype(np.float16)
        np.savez_compressed(OUT / f"{name}.npz", mean=mean, var=var,
                            meta=json.dumps({**meta_common, "scene": kind, "width": w, "height": h, "spp": n, "max_child_rays": depth,
                                             "aspect": aspect, "how": "8 single-threaded runs of render.cpp:152-163 with seeds 1000+7919*i, summed"}))
        print(name, "mean", mean.mean(axis=(0, 1)))


if __name__ == "__main__":
    main()
