"""CPU-side tests: the C-ABI library loads and exports what include/rtw_b200.h declares, the host C++ model mirrors
the reference's API and data (scenes, camera, Config printer, PPM writer, OBJ loader, CLI), and nothing falls back to
a CPU renderer when there is no GPU.  No compute kernels are launched here."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
SUZANNE = str(ROOT / "assets" / "suzanne.obj")


def test_library_exports_every_declared_symbol(rtw):
    header = rtw.HEADER_PATH.read_text()
    declared = set(re.findall(r"RTW_API[^;]*?\b(rtw_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 14
    L = rtw.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in rtw_b200.h but not exported by librtw_b200.so"
    assert declared == set(rtw.ABI), "python binding and header disagree"
    assert L.rtw_abi_version() == int(re.search(r"#define RTW_ABI_VERSION (\d+)", header).group(1))
    exported = subprocess.run(["nm", "-D", "--defined-only", str(rtw.LIB_PATH)], capture_output=True, text=True).stdout
    assert set(re.findall(r" T (rtw_[a-z0-9_]+)", exported)) == declared


def test_struct_layouts_match_the_header(rtw, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "rtw_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(rtw_primitive),sizeof(rtw_material),sizeof(rtw_camera),sizeof(rtw_scene_desc),sizeof(rtw_render_cfg),"
                   "sizeof(rtw_stats),offsetof(rtw_render_cfg,seed),offsetof(rtw_scene_desc,camera));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["/usr/bin/gcc", "-std=c11", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(rtw.Primitive), C.sizeof(rtw.Material), C.sizeof(rtw.Camera), C.sizeof(rtw.SceneDesc), C.sizeof(rtw.RenderCfg),
            C.sizeof(rtw.Stats), rtw.RenderCfg.seed.offset, rtw.SceneDesc.camera.offset]
    assert got == want


def test_no_cpu_fallback(rtw):
    """Without a CUDA device every compute entry point must fail loudly."""
    if rtw.device_count() > 0:
        pytest.skip("a GPU is present")
    scene = rtw.cover_scene(2)
    with pytest.raises(rtw.RtwError, match="no CUDA device|cudaGetDeviceCount"):
        rtw.render(scene, 8, 8, 1)
    with pytest.raises(rtw.RtwError):
        rtw.primary_hits(scene, 8, 8)
    with pytest.raises(rtw.RtwError):
        rtw.DeviceScene(scene)
    r = subprocess.run([str(rtw.EXE_PATH), "-w", "16", "-s", "4"], capture_output=True, text=True)
    assert r.returncode != 0 and "P3" not in r.stdout and "rtw_b200" in r.stderr


def test_product_does_not_reference_the_oracle(rtw):
    """The oracle is test infrastructure: nothing under the package may import, link or execute it."""
    for path in rtw.PKG_DIR.rglob("*"):
        if path.suffix in {".py", ".cu", ".cuh", ".h", ".cpp", ".sh"}:
            text = path.read_text()
            for needle in ("import oracle", "oracle/", "librtw_oracle", "libref_oracle", "rtwo_", "rtw_oracle"):
                assert needle not in text, f"{path} mentions {needle}"


def test_built_library_is_sm_100a_code_with_the_instructions_the_design_claims(rtw):
    """cuobjdump of the in-tree librtw_b200.so (no GPU needed): sm_100a cubins only; bulk-TMA staging + mbarrier (UBLKCP, SYNCS) in
    every kernel that keeps its tables in shared memory; the packed FP32 FMA of sm_100 (FFMA2) in the sphere sweep and nowhere in the
    tree walks (measured slower there); 64-bit reductions for the accumulation; 256-bit global loads for the big-mesh nodes;
    and no tensor-core instruction at all (nothing on this path is a dense contraction)."""
    import collections, re, shutil, subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    elf = subprocess.run([cuobjdump, "-lelf", str(rtw.LIB_PATH)], capture_output=True, text=True).stdout
    cubins = re.findall(r"ELF file\s+\d+: (\S+)", elf)
    assert cubins and all(c.endswith(".sm_100a.cubin") for c in cubins), cubins
    sass = subprocess.run([cuobjdump, "-sass", str(rtw.LIB_PATH)], capture_output=True, text=True).stdout
    per_fn = collections.defaultdict(collections.Counter)
    fn = ""
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        for op in ("FFMA2", "UBLKCP", "SYNCS", "REDG.E.ADD.64", "LDG.E.ENL2.256", "HMMA", "IMMA", "UTCHMMA", "UTCQMMA", "WGMMA"):
            if op in line:
                per_fn[fn][op] += 1
    total = collections.Counter()
    for c in per_fn.values():
        total.update(c)
    assert total["REDG.E.ADD.64"] > 0 and total["LDG.E.ENL2.256"] > 0
    assert not any(total[op] for op in ("HMMA", "IMMA", "UTCHMMA", "UTCQMMA", "WGMMA")), total
    wf = [f for f in per_fn if "k_render_wfILb1E" in f]          # K2w with its tables staged in shared memory
    assert len(wf) >= 8 and all(per_fn[f]["UBLKCP"] > 0 and per_fn[f]["SYNCS"] > 0 and per_fn[f]["FFMA2"] == 0 for f in wf), wf
    assert any("Li32ELi92E" in f for f in wf)                     # the 32-warp, 92-record tier
    sweep2 = [f for f in per_fn if "k_render_sweepILi2E" in f]
    assert sweep2 and all(per_fn[f]["FFMA2"] >= 100 and per_fn[f]["UBLKCP"] > 0 for f in sweep2), {f: dict(per_fn[f]) for f in sweep2}
    assert all(per_fn[f]["FFMA2"] == 0 for f in per_fn if "k_render_bvh" in f)


def test_cover_scene_matches_oracle_scene(rtw, port):
    for nsqrt, moving in [(11, True), (11, False), (3, True)]:
        s = rtw.cover_scene(nsqrt, 1.5, moving)
        pp, pm = port.scene_cover(nsqrt, 1.5, moving).dump()
        assert len(s.prims) == len(pp) and len(s.mats) == len(pm)
        for f in ("kind", "material", "a", "b", "radius"):
            assert np.array_equal(s.prims[f], pp[f]), f
        for f in ("kind", "albedo", "fuzz", "ior"):
            assert np.array_equal(s.mats[f], pm[f]), f
    s = rtw.cover_scene()
    kinds = s.prims["kind"]
    assert len(s.prims) == 485 and (kinds == rtw.RTW_MOVING_SPHERE).sum() == 389 and (kinds == rtw.RTW_SPHERE).sum() == 96
    mk = s.mats["kind"]
    assert [(mk == k).sum() for k in (0, 1, 2)] == [391, 75, 19]


def test_camera_block_matches_oracle(rtw, port):
    import ctypes
    for aspect in (1.5, 1.7777777777777777):
        s = rtw.cover_scene(1, aspect)
        out = (ctypes.c_double * 21)()
        osc = port.scene_cover(1, aspect)
        port.L.rtwo_scene_camera_derived.argtypes = [ctypes.c_void_p, ctypes.c_double * 21]
        port.L.rtwo_scene_camera_derived(osc.h, out)
        c = s.camera
        mine = list(c.origin) + list(c.lower_left) + list(c.horizontal) + list(c.vertical) + list(c.u) + list(c.v) + [c.lens_radius, c.t0, c.t1]
        assert mine == list(out)
    c = rtw.make_camera((1, 0, -1), (0, 0, 0), (0, 1, 0), 35.0, 1.5, 0.01, None, 0, 1)
    assert c.lens_radius == 0.005 and np.isclose(np.linalg.norm(c.horizontal) / np.linalg.norm(c.vertical), 1.5)


def test_obj_loader(rtw, port, tmp_path):
    s = rtw.obj_scene(SUZANNE)
    pp, _ = port.scene_obj(SUZANNE).dump()
    assert len(s.prims) == 968 and np.all(s.prims["kind"] == rtw.RTW_TRIANGLE)
    for f in "abc":
        assert np.array_equal(s.prims[f], pp[f])
    assert len(s.mats) == 1 and s.mats["kind"][0] == rtw.RTW_LAMBERTIAN and np.allclose(s.mats["albedo"][0], 0.5)
    # polygons are fan-triangulated, negative indices are relative, only the first shape is read
    obj = tmp_path / "quad.obj"
    obj.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0.5 0.5 1e0\nf 1 2 3 4\nf -1/1/1 -5//2 -4\ng second\nf 1 2 3\n")
    q = rtw.obj_scene(str(obj))
    assert len(q.prims) == 3
    assert np.array_equal(q.prims["a"][1], [0, 0, 0]) and np.array_equal(q.prims["b"][1], [1, 1, 0]) and np.array_equal(q.prims["c"][1], [0, 1, 0])
    assert np.array_equal(q.prims["a"][2], [0.5, 0.5, 1.0]) and np.array_equal(q.prims["b"][2], [0, 0, 0])
    with pytest.raises(rtw.RtwError, match="Can't load because"):
        rtw.obj_scene(str(tmp_path / "missing.obj"))
    # --all-shapes: every shape of the file, not only shapes[0]
    rtw.host().rtwh_set_obj_all_shapes(1)
    try:
        q4 = rtw.obj_scene(str(obj))
    finally:
        rtw.host().rtwh_set_obj_all_shapes(0)
    assert len(q4.prims) == 4 and np.array_equal(q4.prims["c"][3], [1, 1, 0])


def test_high_poly_stand_in_generator(rtw, tmp_path):
    out = tmp_path / "fine.obj"
    n = C.c_longlong(0)
    assert rtw.host().rtwh_make_mesh(SUZANNE.encode(), str(out).encode(), 2, 7, 0.08, C.byref(n)) == 0
    assert n.value == 968 * 16
    s = rtw.obj_scene(str(out))
    assert len(s.prims) == 968 * 16
    out2 = tmp_path / "fine2.obj"
    rtw.host().rtwh_make_mesh(SUZANNE.encode(), str(out2).encode(), 2, 7, 0.08, None)
    assert out.read_bytes() == out2.read_bytes()  # deterministic


def test_config_printer_and_cli(rtw, oracle_mod):
    buf = C.create_string_buffer(1024)
    n = rtw.host().rtwh_config_string(11, 1.5, 200, 20, 1, 20, 4, buf, 1024)
    want = "Config {\naspect_ratio: 1.5\nnumber_of_balls_sqrt: 11\nmoving_spheres: 1\nimage_width: 200\nsamples_per_pixel: 20\nmax_child_rays: 20\nnthreads: 4\n}\n"
    assert buf.value.decode() == want and n == len(want)
    r = subprocess.run([str(rtw.EXE_PATH), "--dry-run"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == want
    args = ["--dry-run", "-t", "8", "-w1920", "--samples-per-pixel=1024", "-c", "50", "-a", "1.7777777777777777", "-n", "5", "-m", "-q"]
    r = subprocess.run([str(rtw.EXE_PATH), *args], capture_output=True, text=True)
    assert "image_width: 1920\nsamples_per_pixel: 1024\nmax_child_rays: 50\nnthreads: 8" in r.stdout and "aspect_ratio: 1.77778" in r.stdout
    if oracle_mod.ref_available():
        q = subprocess.run([str(oracle_mod.REF_EXE), *args], capture_output=True, text=True)
        assert q.stdout == r.stdout
    r = subprocess.run([str(rtw.EXE_PATH), "--no-such-flag"], capture_output=True, text=True)
    assert r.returncode != 0 and "not expected" in r.stderr
    r = subprocess.run([str(rtw.EXE_PATH), "-w", "abc"], capture_output=True, text=True)
    assert r.returncode != 0


def test_image_geometry_and_effective_spp(rtw):
    assert rtw.image_height(200, 1.5) == 133
    assert rtw.image_height(1920, 1.7777777777777777) == 1080
    assert rtw.image_height(1920, 1.7778) == 1079
    assert rtw.image_height(3840, 1.7777777777777777) == 2160
    H = rtw.host()
    assert H.rtwh_effective_spp(20, 4) == 20 and H.rtwh_effective_spp(20, 8) == 16 and H.rtwh_effective_spp(20, 1) == 20
    assert H.rtwh_effective_spp(3, 8) == 0 and H.rtwh_effective_spp(3, 0) == -1


def test_ppm_writer_matches_write_color(rtw, port, tmp_path):
    rng = np.random.default_rng(1)
    W, H, spp = 37, 11, 20
    acc = np.zeros((H, W, 4), np.float32)
    acc[..., :3] = rng.uniform(0, 1.2, (H, W, 3)) * spp
    acc[0, 0, :3] = [0, spp * 0.25, spp * 4.0]
    acc[..., 3] = spp
    path = tmp_path / "x.ppm"
    assert rtw.host().rtwh_write_ppm(acc.ctypes.data_as(C.c_void_p), W, H, spp, str(path).encode()) == 0
    text = path.read_text()
    assert text.startswith(f"P3\n{W} {H}\n255\n") and text.count("\n") == 3 + W * H
    img = rtw.read_ppm(text)
    want = port.quantize(acc[..., :3].astype(np.float64), spp)
    assert np.array_equal(img, want)
    assert img[0, 0].tolist() == [0, 128, 255]


def test_sample_shard(rtw):
    assert rtw.sample_shard(1024, 0, 8) == (0, 128) and rtw.sample_shard(1024, 7, 8) == (896, 1024)
    assert [rtw.sample_shard(12, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 12)]
    with pytest.raises(ValueError):
        rtw.sample_shard(20, 0, 8)


def test_bvh_kernel_plan(rtw, monkeypatch):
    """Which BVH kernel a scene gets is host logic (plan_bvh, csrc/rtw_internal.h), reported by rtw_flatten_info: the wavefront
    kernel with 32 / 28 / 24 / 20 warps per SM while tables + path records fit in the 227 KB of shared memory, the same kernel with the
    tables in L1/L2 for bigger sphere scenes, the per-lane kernel for meshes; with RTW_MESH_BVH=cw8 scenes with triangles get the
    compressed 8-wide BVH walked by the per-lane state machine, its tables in shared memory while they fit in 72 KB."""
    def plan(scene):
        r = rtw.flatten_info(scene)
        return r["bvh_variant"], r["bvh_warps_per_cta"], r["bvh_tables_in_smem"], r["bvh_smem_bytes"]
    monkeypatch.delenv("RTW_MESH_BVH", raising=False)
    for nsqrt, warps in ((1, 32), (11, 32), (12, 28), (13, 24), (15, 24), (16, 20), (17, 20)):
        v, w, in_smem, smem = plan(rtw.cover_scene(nsqrt))
        assert (v, w, in_smem) == (rtw.BVH_WAVEFRONT, warps, 1), nsqrt
        assert smem <= 227 * 1024 and smem + 4 * 6048 > 227 * 1024 or warps >= 28   # the largest tier that fits
    assert plan(rtw.cover_scene(11))[3] == 46416 + 32 * 5808   # cover scene: 46 416 bytes of tables + 32 warps x 92 records (176 bytes below the cap)
    assert plan(rtw.cover_scene(18))[:3] == (rtw.BVH_WAVEFRONT, 8, 0)
    assert plan(rtw.cover_scene(40))[:3] == (rtw.BVH_WAVEFRONT, 8, 0)
    assert plan(rtw.mesh_on_ground_scene(SUZANNE))[:3] == (rtw.BVH_PERLANE, 8, 0)   # 968 triangles: 108 KB of tables, beyond K2's 72 KB
    cam = dict(lookfrom=(0, 0, 0), lookat=(0, 0, -1), vup=(0, 1, 0), vfov=60.0, aspect=1.0, aperture=0.0, focus_dist=1.0)
    prims = np.zeros(200, rtw.PRIM_DTYPE)
    prims["kind"] = rtw.RTW_TRIANGLE
    rng = np.random.default_rng(0)
    prims["a"] = rng.uniform(-1, 1, (200, 3)); prims["b"] = prims["a"] + 0.1; prims["c"] = prims["a"] + [0.1, 0, 0.05]
    small = rtw.custom_scene(prims, np.zeros(1, rtw.MAT_DTYPE), **cam)
    assert plan(small)[:3] == (rtw.BVH_WAVEFRONT, 32, 1)   # small meshes fit the first tier
    big = np.zeros(4000, rtw.PRIM_DTYPE)
    big["kind"] = rtw.RTW_TRIANGLE
    big["a"] = rng.uniform(-1, 1, (4000, 3)); big["b"] = big["a"] + 0.01; big["c"] = big["a"] + [0.01, 0, 0.005]
    big = rtw.custom_scene(big, np.zeros(1, rtw.MAT_DTYPE), **cam)
    monkeypatch.setenv("RTW_MESH_BVH", "cw8")
    v, w, in_smem, smem = plan(rtw.mesh_on_ground_scene(SUZANNE))
    assert (v, w, in_smem) == (rtw.BVH_CWIDE, 8, 1) and 50_000 < smem < 60_000    # 968 triangles: ~130 wide nodes + 968 records = 57 KB
    assert plan(small)[:3] == (rtw.BVH_CWIDE, 8, 1)
    assert plan(big)[:3] == (rtw.BVH_CWIDE, 8, 0)   # 192 KB of records: read through L1/L2
    assert plan(rtw.cover_scene(11))[:3] == (rtw.BVH_WAVEFRONT, 32, 1)   # sphere-only scenes are not affected
    assert rtw.scene_hash(small) != (monkeypatch.delenv("RTW_MESH_BVH") or rtw.scene_hash(small))   # the cache key knows the tree format


def test_flatten_tables_and_bvh_invariants(rtw, tmp_path, monkeypatch):
    """north_star item 1 on the CPU: the primitive list flattened into sphere / big-sphere / triangle tables and a BVH in
    which every primitive appears exactly once (rtw_flatten_info runs the host half of rtw_scene_upload, no GPU)."""
    r = rtw.flatten_info(rtw.cover_scene())
    assert (r["n_static_spheres"], r["n_moving_spheres"], r["n_big_spheres"], r["n_triangles"]) == (95, 389, 1, 0)
    assert r["n_bvh_nodes"] == 483 and r["leaf_direct"] == 1 and r["bvh_errors"] == 0 and r["bvh_max_depth"] <= 16
    r = rtw.flatten_info(rtw.cover_scene(11, 1.5, False))
    assert (r["n_static_spheres"], r["n_moving_spheres"], r["n_big_spheres"]) == (485, 0, 1)  # 486 static spheres
    # meshes, default: binary tree, triangle records in tree order.  bvh_errors counts primitives referenced zero or several times
    r = rtw.flatten_info(rtw.obj_scene(SUZANNE))
    assert r["n_triangles"] == 968 and r["n_bvh_nodes"] == 967 and r["bvh_errors"] == 0
    r = rtw.flatten_info(rtw.mesh_on_ground_scene(SUZANNE))
    assert r["n_triangles"] == 968 and r["n_big_spheres"] == 1 and r["bvh_errors"] == 0
    for fmt in ("binary", "cw8"):
        # RTW_MESH_BVH=cw8: compressed 8-wide BVH (80-byte nodes); there bvh_errors ALSO counts dequantised child boxes that fail to
        # contain the exact bounds of everything below them (checked bottom-up on the host)
        monkeypatch.setenv("RTW_MESH_BVH", fmt)
        r = rtw.flatten_info(rtw.obj_scene(SUZANNE))
        assert r["n_triangles"] == 968 and r["bvh_errors"] == 0
        if fmt == "cw8":
            assert 968 // 24 <= r["n_bvh_nodes"] <= 967 // 2 and r["bvh_max_depth"] <= 8 and r["bvh_variant"] == rtw.BVH_CWIDE
        # meshes of 1, 2, 3, 7 and 200 triangles, and triangles mixed with small spheres
        cam = dict(lookfrom=(0, 0, 0), lookat=(0, 0, -1), vup=(0, 1, 0), vfov=60.0, aspect=1.0, aperture=0.0, focus_dist=1.0)
        for n, nsph in ((1, 0), (2, 0), (3, 0), (7, 0), (200, 0), (1, 1), (50, 30), (5, 400)):
            rng = np.random.default_rng(100 + n + nsph)
            prims = np.zeros(n + nsph, rtw.PRIM_DTYPE)
            prims["kind"][:n] = rtw.RTW_TRIANGLE
            prims["a"] = rng.uniform(-3, 3, (n + nsph, 3)); prims["b"] = prims["a"] + rng.uniform(-0.3, 0.3, (n + nsph, 3)); prims["c"] = prims["a"] + rng.uniform(-0.3, 0.3, (n + nsph, 3))
            prims["kind"][n:] = rng.integers(0, 2, nsph)   # static and moving spheres
            prims["radius"][n:] = rng.uniform(0.05, 0.4, nsph)
            r = rtw.flatten_info(rtw.custom_scene(prims, np.zeros(1, rtw.MAT_DTYPE), **cam))
            assert r["bvh_errors"] == 0 and r["n_triangles"] == n, (fmt, n, nsph, r)
            assert 1 <= r["n_bvh_nodes"] <= max(n + nsph - 1, 1)
        # degenerate mesh: 3000 copies of one triangle (coincident centroids) and a sliver spanning 60 orders of magnitude
        prims = np.zeros(3000, rtw.PRIM_DTYPE); prims["kind"] = rtw.RTW_TRIANGLE; prims["b"] = [1, 0, 0]; prims["c"] = [0, 1, 0]
        r = rtw.flatten_info(rtw.custom_scene(prims, np.zeros(1, rtw.MAT_DTYPE), **cam))
        assert r["bvh_errors"] == 0 and r["bvh_max_depth"] <= (32 if fmt == "cw8" else 64)
        prims = np.zeros(1500, rtw.PRIM_DTYPE); prims["kind"] = rtw.RTW_TRIANGLE
        prims["a"][:, 0] = 1e-30 * 1.08 ** np.arange(1500); prims["b"] = prims["a"] * 1.01; prims["c"] = prims["a"] + [0, 1e-33, 0]
        r = rtw.flatten_info(rtw.custom_scene(prims, np.zeros(1, rtw.MAT_DTYPE), **cam))
        assert r["bvh_errors"] == 0 and r["bvh_max_depth"] <= (32 if fmt == "cw8" else 64)
    monkeypatch.delenv("RTW_MESH_BVH")
    cam = dict(lookfrom=(0, 0, 0), lookat=(0, 0, -1), vup=(0, 1, 0), vfov=60.0, aspect=1.0, aperture=0.0, focus_dist=1.0)
    mats = np.zeros(1, rtw.MAT_DTYPE)
    for n in (0, 1, 2, 3, 9, 100):
        prims = np.zeros(n, rtw.PRIM_DTYPE)
        prims["a"] = prims["b"] = np.random.default_rng(n).uniform(-5, 5, (n, 3))
        prims["radius"] = 0.3
        prims["kind"][: n // 2] = rtw.RTW_MOVING_SPHERE
        prims["b"][: n // 2] += 0.5
        r = rtw.flatten_info(rtw.custom_scene(prims, mats, **cam))
        assert r["bvh_errors"] == 0 and r["n_bvh_nodes"] == max(n - 1, 1 if n else 0)
        assert r["n_static_spheres"] + r["n_moving_spheres"] == n and r["n_moving_spheres"] == n // 2
    # coincident primitives (all centroids equal) must still build a finite tree
    prims = np.zeros(50, rtw.PRIM_DTYPE); prims["radius"] = 1.0
    r = rtw.flatten_info(rtw.custom_scene(prims, mats, **cam))
    assert r["bvh_errors"] == 0 and r["bvh_max_depth"] <= 50
    # centroids spread over 60 orders of magnitude: plain SAH peels a few primitives per level (57 levels measured); the builder
    # switches to median splits below level 32, so the tree stays inside the kernels' 64-entry traversal stack
    n = 1500
    prims = np.zeros(n, rtw.PRIM_DTYPE)
    prims["a"][:, 0] = 1e-30 * 1.08 ** np.arange(n); prims["b"] = prims["a"]; prims["radius"] = 1e-37
    r = rtw.flatten_info(rtw.custom_scene(prims, mats, **cam))
    assert r["bvh_errors"] == 0 and 32 < r["bvh_max_depth"] <= 32 + 11
    # the multi-primitive leaf layout (tuning knob) keeps the same invariants
    import os, subprocess, sys, textwrap
    code = textwrap.dedent(f"""
        import importlib, sys
        sys.path.insert(0, {str(ROOT)!r})
        rtw = importlib.import_module("raytracing-one-weekend_b200")
        r = rtw.flatten_info(rtw.cover_scene())
        assert r["leaf_direct"] == 0 and r["bvh_errors"] == 0 and r["n_bvh_nodes"] < 300, r
    """)
    subprocess.run([sys.executable, "-c", code], check=True, env={**os.environ, "RTW_BVH_LEAF": "4"})
    # errors
    bad = rtw.cover_scene(1)
    bad.prims["material"][0] = 99
    with pytest.raises(rtw.RtwError, match="material"):
        rtw.flatten_info(bad)
    bad = rtw.cover_scene(1)
    bad.prims["kind"][0] = 7
    with pytest.raises(rtw.RtwError, match="primitive kind"):
        rtw.flatten_info(bad)


def test_variant_and_oo_models_flatten_identically(rtw):
    """north_star keeps both primitive containers (variant-primitives.h / oo-primitives.h behind primitive-model.h): the host side
    built with -DRTWEEKEND_USE_VARIANT_PRIMITIVES hands the same arrays to the C ABI as the default (virtual) build, scene by scene."""
    oo, var = rtw.host(), rtw.host_variant()
    assert oo.rtwh_primitive_model() == b"oo" and var.rtwh_primitive_model() == b"variant"
    for args in [(11, 1.5, True), (11, 1.7777777777777777, False), (3, 1.5, True)]:
        a, b = rtw.cover_scene(*args, H=oo), rtw.cover_scene(*args, H=var)
        assert a.prims.tobytes() == b.prims.tobytes() and a.mats.tobytes() == b.mats.tobytes()
        assert bytes(a.camera) == bytes(b.camera)
        assert rtw.scene_hash(a) == rtw.scene_hash(b)
    a, b = rtw.obj_scene(SUZANNE, H=oo), rtw.obj_scene(SUZANNE, H=var)
    assert len(a.prims) == 968 and a.prims.tobytes() == b.prims.tobytes() and a.mats.tobytes() == b.mats.tobytes()
    # the executables of both builds parse the same command line
    out = [subprocess.run([str(e), "--dry-run", "-w", "320", "-s", "7"], capture_output=True, text=True) for e in (rtw.EXE_PATH, rtw.EXE_VARIANT_PATH)]
    assert out[0].returncode == 0 and out[0].stdout == out[1].stdout and "image_width: 320" in out[0].stdout


def test_scene_hash_follows_the_content(rtw):
    """rtw_scene_hash keys the device scene cache and binds checkpoints to their scene: equal content, equal hash; any changed
    double, material, camera field or primitive count, a different one."""
    s = rtw.cover_scene()
    h = rtw.scene_hash(s)
    assert h == rtw.scene_hash(rtw.cover_scene()) and h != 0
    seen = {h}
    for mutate in (lambda q: q.prims["a"].__setitem__((17, 1), q.prims["a"][17, 1] + 1e-12),
                   lambda q: q.prims["radius"].__setitem__(484, 1.0000001),
                   lambda q: q.mats["albedo"].__setitem__((3, 2), 0.25),
                   lambda q: q.prims["material"].__setitem__(5, 6),
                   lambda q: setattr(q.camera, "t1", 0.5)):
        q = rtw.cover_scene()
        mutate(q)
        seen.add(rtw.scene_hash(q))
    q = rtw.cover_scene()
    q.prims = q.prims[:-1].copy()
    seen.add(rtw.scene_hash(q))
    assert len(seen) == 7
    big = rtw.cover_scene(120)   # 57 603 primitives: the chunked (multi-threaded) hashing path
    assert len(big.prims) * big.prims.itemsize > (4 << 20)
    hb = rtw.scene_hash(big)
    assert hb == rtw.scene_hash(rtw.cover_scene(120))
    big.prims["b"][40000, 2] += 1e-9
    assert rtw.scene_hash(big) != hb
