"""Pins the CPU oracle (oracle/rtw_oracle.c, the plain-C restatement) to the reference.

Golden fixtures under tests/golden/ were produced by tests/golden/make_golden.py from the UNMODIFIED reference sources
(oracle/_ref).  Where oracle/_ref is present (build container, and the GPU box through the prebuilt files) the port is
additionally compared with the reference live.
"""
import hashlib

import numpy as np
import pytest

SUZANNE = str(__import__("pathlib").Path(__file__).resolve().parents[1] / "assets" / "suzanne.obj")


def test_host_rng_matches_libstdcxx(port):
    # first three std::uniform_real_distribution<double>(0,1)(std::mt19937{}) values of libstdc++
    port.seed(5489)
    got = [port.random_double() for _ in range(3)]
    assert got == [0.1354770042967805, 0.8350085899945795, 0.96886777112423139]


def test_philox_known_answers(port):
    # Random123 known-answer vectors for philox4x32-10
    assert port.philox([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert port.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert port.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_port_reproduces_reference_image_bit_for_bit(port, golden):
    """The reference's `rtweekend -t 1` output (default config) is bit-reproducible; the port must print the same P3."""
    g = golden("cover_default_t1.npz")
    W, H, spp, depth = g["meta"]["width"], g["meta"]["height"], g["meta"]["spp"], g["meta"]["max_child_rays"]
    sc = port.scene_cover(11, 1.5, True, seed=5489)  # scene consumes the default-seeded stream, render continues on it
    s, _, rays = sc.render_linear(W, H, spp, depth, seed=None, want_sumsq=False)
    rgb = port.quantize(s, spp)
    assert np.array_equal(rgb, g["rgb"])
    text = f"P3\n{W} {H}\n255\n" + "".join(f"{p[0]} {p[1]} {p[2]}\n" for p in rgb.reshape(-1, 3))
    assert hashlib.md5(text.encode()).hexdigest() == g["meta"]["md5_of_p3_text"] == "97d9c29de0118c11bd767191ff4e1d4a"
    assert abs(rays / (W * H * spp) - 2.30) < 0.01  # rays per path measured by the survey


@pytest.mark.parametrize("name,time", [("cover_primary_200x133_t0.npz", 0.0), ("cover_primary_200x133_t0.5.npz", 0.5)])
def test_port_primary_hits_cover(port, golden, name, time):
    g = golden(name)
    pid, t, nrm, front = port.scene_cover().primary_hits(200, 133, time)
    assert np.array_equal(pid, g["id"])
    assert np.array_equal(t, g["t"])  # same doubles: same formulas, same operation order
    assert np.array_equal(nrm.astype(np.float32), g["normal"])
    assert np.array_equal(front, g["front"])


def test_port_primary_hits_suzanne(port, golden):
    g = golden("suzanne_primary_200x133.npz")
    pid, t, nrm, front = port.scene_obj(SUZANNE).primary_hits(200, 133, 0.0)
    assert np.array_equal(pid, g["id"])
    assert np.array_equal(t, g["t"])
    assert np.array_equal(nrm.astype(np.float32), g["normal"])
    assert (pid >= 0).mean() > 0.9


def _z_stats(mean_a, var_a, n_a, mean_b, var_b, n_b):
    """Per-channel z-score of the image-mean difference and per-pixel z-scores."""
    se_pix = np.sqrt(var_a / n_a + var_b / n_b)
    npix = mean_a.shape[0] * mean_a.shape[1]
    z_img = (mean_a - mean_b).mean(axis=(0, 1)) / (np.sqrt((se_pix ** 2).sum(axis=(0, 1))) / npix)
    with np.errstate(divide="ignore", invalid="ignore"):
        z_pix = np.where(se_pix > 0, (mean_a - mean_b) / se_pix, 0.0)
    return z_img, z_pix


def test_port_philox_contract_matches_reference_distribution(port, golden):
    """The sampling contract of the new renderer (Philox + direct inversion), evaluated in double by the port, must
    converge to the reference's own converged render: per-channel image mean within 3 sigma of Monte Carlo noise."""
    g = golden("cover_converged_120x80.npz")
    m = g["meta"]
    sc = port.scene_cover(11, m["aspect"], True)
    spp = 96
    s, q, _ = port.render_philox(sc, m["width"], m["height"], 0, spp, m["max_child_rays"], seed=11, nthreads=8, want_sumsq=True)
    mean = s / spp
    var = np.maximum(q / spp - mean ** 2, 0) * spp / (spp - 1)
    z_img, z_pix = _z_stats(mean, var, spp, g["mean"].astype(np.float64), g["var"].astype(np.float64), m["spp"])
    assert np.all(np.abs(z_img) < 3.0), z_img
    assert (np.abs(z_pix) > 5).mean() < 2e-3


def test_port_matches_reference_live(port, oracle_mod):
    """Where the compiled reference is present: other seeds, sizes and the static-sphere scene, bit for bit."""
    if not oracle_mod.ref_available():
        pytest.skip("oracle/_ref not built here")
    ref = oracle_mod.ref()
    for moving, aspect, W, H, spp, depth, seed in [(True, 1.5, 40, 26, 6, 20, 1), (False, 1.7777777777777777, 48, 27, 4, 50, 99),
                                                   (True, 1.5, 30, 20, 3, 0, 5), (True, 1.5, 30, 20, 3, 1, 6)]:
        a, qa, _ = ref.scene_cover(11, aspect, moving).render_linear(W, H, spp, depth, seed=seed)
        b, qb, _ = port.scene_cover(11, aspect, moving).render_linear(W, H, spp, depth, seed=seed)
        assert np.array_equal(a, b) and np.array_equal(qa, qb)
    a, _, _ = ref.scene_obj(SUZANNE).render_linear(24, 16, 2, 20, seed=3)
    b, _, _ = port.scene_obj(SUZANNE).render_linear(24, 16, 2, 20, seed=3)
    assert np.array_equal(a, b)
    # scenes, including the constructor-argument evaluation order of g++ (main.cpp:46)
    for moving in (True, False):
        rp, rm = ref.scene_cover(5, 1.5, moving).dump()
        pp, pm = port.scene_cover(5, 1.5, moving).dump()
        for f in ("kind", "material", "a", "b", "radius"):
            assert np.array_equal(rp[f], pp[f])
        for f in ("kind", "albedo", "fuzz", "ior"):
            assert np.array_equal(rm[f], pm[f])


# ---- closed-form unit vectors for the formulas on the path (SURVEY 8(c) pin 4) ---------------------------------------
def test_sphere_hit_closed_form(port):
    # ray along +z from the origin with |d| = 2 towards a unit sphere at z = 5: roots at distance 4 and 6 -> t = 2, 3
    t, p, n, front = port.hit_sphere([0, 0, 0], [0, 0, 2], 0.001, np.inf, [0, 0, 5], 1.0)
    assert t == 2.0 and np.allclose(p, [0, 0, 4]) and np.allclose(n, [0, 0, -1]) and front
    # origin inside: the nearer root is negative, the farther one is taken; normal flipped against the ray
    t, p, n, front = port.hit_sphere([0, 0, 5], [0, 0, 2], 0.001, np.inf, [0, 0, 5], 1.0)
    assert t == 0.5 and np.allclose(n, [0, 0, -1]) and not front
    # negative radius flips front_facing (common-model.cpp:88)
    t, p, n, front = port.hit_sphere([0, 0, 0], [0, 0, 2], 0.001, np.inf, [0, 0, 5], -1.0)
    assert t == 2.0 and not front and np.allclose(n, [0, 0, 1])
    # tmax below the near root -> miss; tmin above near root but below far root -> far root
    assert port.hit_sphere([0, 0, 0], [0, 0, 2], 0.001, 1.9, [0, 0, 5], 1.0) is None
    t, *_ = port.hit_sphere([0, 0, 0], [0, 0, 2], 2.5, np.inf, [0, 0, 5], 1.0)
    assert t == 3.0
    assert port.hit_sphere([0, 0, 0], [0, 2, 0], 0.001, np.inf, [0, 0, 5], 1.0) is None


def test_triangle_hit_closed_form(port):
    a, b, c = [0, 0, 1], [1, 0, 1], [0, 1, 1]
    # n = e1 x e2 = (0,0,1); det = -d.n must be >= 1e-6: only rays travelling towards -z hit (back-face culling, Q7)
    assert port.hit_triangle([0.2, 0.2, 0], [0, 0, 1], 0.001, np.inf, a, b, c) is None
    t, p, n = port.hit_triangle([0.2, 0.2, 3], [0, 0, -4], 0.001, np.inf, a, b, c)
    assert t == 0.5 and np.allclose(p, [0.2, 0.2, 1]) and np.array_equal(n, [0, 0, 1])  # un-normalised e1 x e2
    t, p, n = port.hit_triangle([0.2, 0.2, 3], [0, 0, -4], 0.001, np.inf, [0, 0, 1], [2, 0, 1], [0, 2, 1])
    assert np.array_equal(n, [0, 0, 4])
    assert port.hit_triangle([0.8, 0.8, 3], [0, 0, -4], 0.001, np.inf, a, b, c) is None  # u + v > 1
    # scale-dependent determinant threshold: a tiny triangle seen with a short direction vector is culled
    s = 0.9e-3  # |e1 x e2| = 0.81e-6
    assert port.hit_triangle([0.2 * s, 0.2 * s, 1], [0, 0, -1], 0.001, np.inf, [0, 0, 0], [s, 0, 0], [0, s, 0]) is None  # det = 0.81e-6 < 1e-6
    assert port.hit_triangle([0.2 * s, 0.2 * s, 1], [0, 0, -10], 0.001, np.inf, [0, 0, 0], [s, 0, 0], [0, s, 0]) is not None


def _mat(kind, albedo=(1, 1, 1), fuzz=0.0, ior=0.0):
    return (kind, 0, albedo, fuzz, ior)


def test_scatter_closed_form(port):
    n = np.array([0.0, 1.0, 0.0])
    ball = np.array([0.1, 0.2, 0.3])
    d, att = port.scatter(_mat(0, (0.2, 0.4, 0.6)), [1, -1, 0], n, True, ball, 0.5)
    assert np.allclose(d, n + ball) and np.allclose(att, [0.2, 0.4, 0.6])  # no normalisation (Q2)
    d, att = port.scatter(_mat(1, (0.9, 0.8, 0.7), fuzz=0.5), [3, -4, 0], n, True, ball, 0.5)
    assert np.allclose(d, np.array([3, 4, 0]) + 0.5 * ball)  # keeps |d_in|, always scatters (Q3)
    d, _ = port.scatter(_mat(1, fuzz=7.0), [3, -4, 0], n, True, ball, 0.5)
    assert np.allclose(d, np.array([3, 4, 0]) + 1.0 * ball)  # fuzz clamped to 1
    # dielectric, normal incidence from outside, ior 1.5: reflectance r0 = 0.04
    d, att = port.scatter(_mat(2, ior=1.5), [0, -2, 0], n, True, [0, 0, 0], 0.5)
    assert np.allclose(d, [0, -1, 0]) and np.allclose(att, [1, 1, 1])  # refracted, unit length
    d, _ = port.scatter(_mat(2, ior=1.5), [0, -2, 0], n, True, [0, 0, 0], 0.03)
    assert np.allclose(d, [0, 1, 0])  # coin below reflectance -> reflected
    # total internal reflection from inside at 60 degrees (sin = .866 * 1.5 > 1)
    din = np.array([np.sin(np.pi / 3), -np.cos(np.pi / 3), 0])
    d, _ = port.scatter(_mat(2, ior=1.5), din, n, False, [0, 0, 0], 0.999)
    assert np.allclose(d, [din[0], -din[1], 0])
    # Snell from outside at 45 degrees
    din = np.array([1, -1, 0]) / np.sqrt(2)
    d, _ = port.scatter(_mat(2, ior=1.5), din * 3, n, True, [0, 0, 0], 0.999)
    assert np.isclose(d[0], np.sin(np.pi / 4) / 1.5) and np.isclose(np.linalg.norm(d), 1.0)


def test_sky_and_quantise(port):
    assert np.allclose(port.sky([0, 5, 0]), [0.5, 0.7, 1.0])
    assert np.allclose(port.sky([0, -5, 0]), [1.0, 1.0, 1.0])
    assert np.allclose(port.sky([3, 0, 0]), [0.75, 0.85, 1.0])
    sums = np.array([[[0.0, 20.0 * 0.25, 20.0 * 1.0], [20.0 * 4.0, 20 * 0.999 ** 2, 20 * 0.5]]])
    q = port.quantize(sums, 20)
    assert q.tolist() == [[[0, 128, 255], [255, 255, int(256 * np.sqrt(0.5))]]]
