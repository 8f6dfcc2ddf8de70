#!/usr/bin/env python
"""One-off CPU study (needs oracle/_ref, i.e. the container that has /root/reference): is the renderer's sampling contract
(Philox stream + direct inversion, evaluated in double by the oracle port) unbiased against the reference's own sampler?
BASELINE config 2's scene, camera and depth at 384x216; R reference realisations of 1024 spp (8 single-threaded runs of
render.cpp:152-163 each, distinct seeds) against P port renders of 1024 spp (distinct Philox keys).  Prints the relative
difference of the image mean per channel and its z score.  Takes about (1 + 3 P/R) x R minutes on 8 cores.
    python scripts/contract_bias_study.py [R=6] [P=4]
Round-1 result (R=6, P=4): relative difference (-1.4, +1.6, -0.4)e-5, z (-0.8, +0.9, -0.2)."""
import multiprocessing as mp
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))


def main():
    import make_golden as mg
    import oracle
    oracle.build()
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    kind, w, h, depth, aspect = "cover", 384, 216, 50, 1.7777777777777777
    ref_sum, ref_sq, ref_n = 0.0, 0.0, 0
    for rep in range(R):
        jobs = [(kind, w, h, 128, depth, 9000000 + 15485863 * rep + 104729 * i, aspect) for i in range(8)]
        with mp.Pool(8) as pool:
            res = pool.map(mg._linear_worker, jobs)
        ref_sum = ref_sum + sum(r[0] for r in res); ref_sq = ref_sq + sum(r[1] for r in res); ref_n += 1024
        print("reference spp", ref_n, "mean", (ref_sum / ref_n).mean(axis=(0, 1)), flush=True)
    ref_mean = ref_sum / ref_n
    var = np.maximum(ref_sq / ref_n - ref_mean * ref_mean, 0.0)
    port = oracle.port()
    osc = port.scene_cover(11, aspect, True)
    port_sum, port_n = 0.0, 0
    for k in range(P):
        s, _, _ = port.render_philox(osc, w, h, 0, 1024, depth, seed=31337 + 7919 * k, nthreads=8)
        port_sum = port_sum + s; port_n += 1024
        print("port spp", port_n, "mean", (port_sum / port_n).mean(axis=(0, 1)), flush=True)
    d = (port_sum / port_n - ref_mean).mean(axis=(0, 1))
    se = np.sqrt((var / ref_n + var / port_n).mean(axis=(0, 1)) / (w * h))
    print("difference of the image mean", d, "se", se, "z", d / se, "relative", d / ref_mean.mean(axis=(0, 1)))


if __name__ == "__main__":
    main()
