"""raytracing-one-weekend_b200 -- B200-native renderer for the hot path of joaotavora/raytracing-one-weekend.

The product is native: ``librtw_b200.so`` (hand-written sm_100a CUDA behind the C ABI of ``include/rtw_b200.h``) and
``librtweekend_host.so`` / ``rtweekend`` (the C++ host side mirroring the reference's ``render.h`` API).  This module
is a thin ``ctypes`` binding over both for the tests and ``bench.py``; it contains no rendering logic and no CPU
fallback: every compute call raises ``RtwError`` when the CUDA library cannot do the work.

Reference interfaces mirrored (all in /root/reference/src): ``Config`` render.h:11-20, ``render`` render.h:35,
``lots_of_balls`` / ``foo`` main.cpp:23-136.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
LIB_PATH = PKG_DIR / "librtw_b200.so"
HOST_LIB_PATH = PKG_DIR / "librtweekend_host.so"
HOST_VARIANT_LIB_PATH = PKG_DIR / "librtweekend_host_variant.so"   # built with -DRTWEEKEND_USE_VARIANT_PRIMITIVES
EXE_PATH = PKG_DIR / "rtweekend"
EXE_VARIANT_PATH = PKG_DIR / "rtweekend_variant"
HEADER_PATH = REPO_ROOT / "include" / "rtw_b200.h"

RTW_SPHERE, RTW_MOVING_SPHERE, RTW_TRIANGLE = 0, 1, 2
RTW_LAMBERTIAN, RTW_METAL, RTW_DIELECTRIC = 0, 1, 2
KERNEL_AUTO, KERNEL_SPHERES_SMEM, KERNEL_BVH, KERNEL_BVH_PERLANE = 0, 1, 2, 3
BVH_NONE, BVH_PERLANE, BVH_WAVEFRONT, BVH_CWIDE = 0, 1, 2, 3
FLAG_STATS = 1
FLAG_SPLIT_ROWS = 2
FLAG_NO_SCENE_CACHE = 4
FLAG_BVH_BUILD_GPU = 16
FLAG_BVH_BUILD_HOST = 32


class RtwError(RuntimeError):
    pass


# ----------------------------------------------------------------------------------------------------------
# C structs of include/rtw_b200.h
# ----------------------------------------------------------------------------------------------------------
class Primitive(C.Structure):
    _fields_ = [("kind", C.c_int32), ("material", C.c_int32), ("a", C.c_double * 3), ("b", C.c_double * 3),
                ("c", C.c_double * 3), ("radius", C.c_double)]


class Material(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("albedo", C.c_double * 3), ("fuzz", C.c_double),
                ("ior", C.c_double)]


class Camera(C.Structure):
    _fields_ = [("origin", C.c_double * 3), ("lower_left", C.c_double * 3), ("horizontal", C.c_double * 3),
                ("vertical", C.c_double * 3), ("u", C.c_double * 3), ("v", C.c_double * 3),
                ("lens_radius", C.c_double), ("t0", C.c_double), ("t1", C.c_double)]


class SceneDesc(C.Structure):
    _fields_ = [("prims", C.POINTER(Primitive)), ("nprims", C.c_int64), ("mats", C.POINTER(Material)),
                ("nmats", C.c_int64), ("camera", Camera)]


class RenderCfg(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("sample_begin", C.c_int32), ("sample_end", C.c_int32),
                ("max_child_rays", C.c_int32), ("kernel", C.c_int32), ("seed", C.c_uint64), ("device", C.c_int32),
                ("flags", C.c_int32), ("rays_per_lane", C.c_int32), ("row_tile_rows", C.c_int32), ("row_tile_count", C.c_int32),
                ("row_tile_index", C.c_int32)]


class FlattenReport(C.Structure):
    _fields_ = [("n_static_spheres", C.c_int32), ("n_moving_spheres", C.c_int32), ("n_big_spheres", C.c_int32), ("n_triangles", C.c_int32),
                ("n_bvh_nodes", C.c_int32), ("bvh_max_depth", C.c_int32), ("leaf_direct", C.c_int32), ("reserved", C.c_int32),
                ("arena_bytes", C.c_int64), ("bvh_errors", C.c_int64), ("flatten_ms", C.c_double), ("bvh_build_ms", C.c_double),
                ("bvh_variant", C.c_int32), ("bvh_warps_per_cta", C.c_int32), ("bvh_tables_in_smem", C.c_int32), ("reserved2", C.c_int32),
                ("bvh_smem_bytes", C.c_int64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Stats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("rays", C.c_uint64), ("sphere_tests", C.c_uint64),
                ("sphere_candidates", C.c_uint64), ("tri_tests", C.c_uint64), ("node_visits", C.c_uint64),
                ("kernel_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double), ("total_ms", C.c_double),
                ("bvh_build_gpu_ms", C.c_double), ("kernel_used", C.c_int32), ("launches", C.c_int32), ("bvh_variant", C.c_int32), ("scene_cache_hit", C.c_int32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


PRIM_DTYPE = np.dtype([("kind", "<i4"), ("material", "<i4"), ("a", "<f8", 3), ("b", "<f8", 3), ("c", "<f8", 3),
                       ("radius", "<f8")])
MAT_DTYPE = np.dtype([("kind", "<i4"), ("reserved", "<i4"), ("albedo", "<f8", 3), ("fuzz", "<f8"), ("ior", "<f8")])
assert PRIM_DTYPE.itemsize == C.sizeof(Primitive) == 88
assert MAT_DTYPE.itemsize == C.sizeof(Material) == 48

# Entry points declared in include/rtw_b200.h: name -> (restype, argtypes)
_VP = C.c_void_p
ABI = {
    "rtw_abi_version": (C.c_int, []),
    "rtw_last_error": (C.c_char_p, []),
    "rtw_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "rtw_scene_upload": (C.c_int, [C.POINTER(SceneDesc), C.c_int32, C.POINTER(_VP)]),
    "rtw_scene_upload_ex": (C.c_int, [C.POINTER(SceneDesc), C.c_int32, C.c_int32, C.POINTER(_VP)]),
    "rtw_scene_check": (C.c_int, [_VP, C.POINTER(FlattenReport)]),
    "rtw_scene_free": (None, [_VP]),
    "rtw_scene_update": (C.c_int, [_VP, C.POINTER(SceneDesc)]),
    "rtw_kernel_launches": (C.c_ulonglong, []),
    "rtw_render": (C.c_int, [C.POINTER(SceneDesc), C.POINTER(RenderCfg), _VP, C.POINTER(Stats)]),
    "rtw_render_rgb8": (C.c_int, [C.POINTER(SceneDesc), C.POINTER(RenderCfg), _VP, C.POINTER(Stats)]),
    "rtw_prewarm": (C.c_int, [C.c_int32, C.c_int32]),
    "rtw_release_cached_buffers": (None, []),
    "rtw_scene_hash": (C.c_int, [C.POINTER(SceneDesc), C.POINTER(C.c_uint64)]),
    "rtw_render_device": (C.c_int, [_VP, C.POINTER(RenderCfg), _VP, _VP, C.POINTER(Stats)]),
    "rtw_accum_to_float": (C.c_int, [_VP, _VP, C.c_int64, C.c_int32, _VP]),
    "rtw_row_tile_local_rows": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32]),
    "rtw_untile_accum": (C.c_int, [_VP, _VP, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _VP]),
    "rtw_render_multi_gpu": (C.c_int, [C.POINTER(SceneDesc), C.POINTER(RenderCfg), C.c_int32, _VP, C.POINTER(Stats)]),
    "rtw_render_multi_gpu_rgb8": (C.c_int, [C.POINTER(SceneDesc), C.POINTER(RenderCfg), C.c_int32, _VP, C.POINTER(Stats)]),
    "rtw_finalize_rgb8": (C.c_int, [_VP, C.c_int64, C.c_int32, C.c_int32, _VP]),
    "rtw_finalize_rgb8_device": (C.c_int, [_VP, C.c_int64, C.c_int32, C.c_int32, _VP, _VP]),
    "rtw_primary_hits": (C.c_int, [C.POINTER(SceneDesc), C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_int32,
                                   C.c_int32, _VP, _VP, _VP, _VP]),
    "rtw_debug_scatter": (C.c_int, [C.c_int32, C.c_int64, C.POINTER(Material), _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "rtw_debug_samples": (C.c_int, [C.c_int32, C.c_int64, C.c_uint64, _VP, _VP, _VP]),
    "rtw_flatten_info": (C.c_int, [C.POINTER(SceneDesc), C.POINTER(FlattenReport)]),
    "rtw_fp32_peak": (C.c_int, [C.c_int32, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

_lib = None
_host = None
_host_variant = None


def build(force: bool = False) -> None:
    """Compile the CUDA library (nvcc, sm_100a) and the C++ host side in tree.  No-op when up to date."""
    args = ["-f"] if force else []
    for script in (PKG_DIR / "csrc" / "build.sh", PKG_DIR / "host" / "build.sh"):
        r = subprocess.run(["bash", str(script), *args], capture_output=True, text=True)
        if r.returncode != 0:
            raise RtwError(f"{script} failed:\n{r.stdout}\n{r.stderr}")


def lib() -> C.CDLL:
    """librtw_b200.so with typed entry points.  Raises when the library is missing: there is no fallback."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RtwError(f"{LIB_PATH} is missing: run __graft_entry__.build() (nvcc -gencode arch=compute_100a,code=sm_100a)")
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in ABI.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _load_host(path: Path) -> C.CDLL:
    lib()  # dependency
    if not path.exists():
        raise RtwError(f"{path} is missing: run __graft_entry__.build()")
    H = C.CDLL(str(path))
    H.rtwh_last_error.restype = C.c_char_p
    H.rtwh_random_double.restype = C.c_double
    H.rtwh_seed.argtypes = [C.c_uint]
    H.rtwh_scene_cover.restype = _VP
    H.rtwh_scene_cover.argtypes = [C.c_int, C.c_double, C.c_int]
    H.rtwh_scene_obj.restype = _VP
    H.rtwh_scene_obj.argtypes = [C.c_char_p, C.c_double]
    H.rtwh_scene_mesh_on_ground.restype = _VP
    H.rtwh_scene_mesh_on_ground.argtypes = [C.c_char_p, C.c_double]
    H.rtwh_scene_free.argtypes = [_VP]
    H.rtwh_scene_nprims.restype = C.c_longlong
    H.rtwh_scene_nprims.argtypes = [_VP]
    H.rtwh_scene_nmats.restype = C.c_longlong
    H.rtwh_scene_nmats.argtypes = [_VP]
    H.rtwh_scene_flatten.argtypes = [_VP, _VP, _VP, C.POINTER(SceneDesc)]
    H.rtwh_camera.argtypes = [C.c_double * 3, C.c_double * 3, C.c_double * 3, C.c_double, C.c_double, C.c_double,
                              C.c_double, C.c_double, C.c_double, C.POINTER(Camera)]
    H.rtwh_render_to_file.argtypes = [_VP, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int,
                                      C.c_ulonglong, C.c_int]
    H.rtwh_config_string.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
    H.rtwh_image_height.argtypes = [C.c_int, C.c_double]
    H.rtwh_effective_spp.argtypes = [C.c_int, C.c_int]
    H.rtwh_make_mesh.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_uint, C.c_double, C.POINTER(C.c_longlong)]
    H.rtwh_write_ppm.argtypes = [_VP, C.c_int, C.c_int, C.c_int, C.c_char_p]
    H.rtwh_primitive_model.restype = C.c_char_p
    return H


def host() -> C.CDLL:
    """librtweekend_host.so: the product's own scene builders / flatten / render() (virtual primitive model)."""
    global _host
    if _host is None:
        _host = _load_host(HOST_LIB_PATH)
    return _host


def host_variant() -> C.CDLL:
    """The same host library compiled with -DRTWEEKEND_USE_VARIANT_PRIMITIVES (std::variant primitive store)."""
    global _host_variant
    if _host_variant is None:
        _host_variant = _load_host(HOST_VARIANT_LIB_PATH)
    return _host_variant


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise RtwError(f"{what} failed ({rc}): {lib().rtw_last_error().decode()}")


def kernel_launches() -> int:
    """Kernels launched by librtw_b200.so in this process so far (counted at every launch site)."""
    return int(lib().rtw_kernel_launches())


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().rtw_device_count(C.byref(n))
    return n.value if rc == 0 else 0


# ----------------------------------------------------------------------------------------------------------
# Scenes
# ----------------------------------------------------------------------------------------------------------
@dataclass
class Scene:
    """A flattened scene: numpy record arrays laid out as rtw_primitive / rtw_material + the camera block."""
    prims: np.ndarray
    mats: np.ndarray
    camera: Camera
    params: dict | None = None  # camera construction parameters when known (for the oracle)

    def desc(self) -> SceneDesc:
        d = SceneDesc()
        self.prims = np.ascontiguousarray(self.prims, dtype=PRIM_DTYPE)
        self.mats = np.ascontiguousarray(self.mats, dtype=MAT_DTYPE)
        d.prims = self.prims.ctypes.data_as(C.POINTER(Primitive))
        d.nprims = len(self.prims)
        d.mats = self.mats.ctypes.data_as(C.POINTER(Material))
        d.nmats = len(self.mats)
        d.camera = self.camera
        return d


def make_camera(lookfrom, lookat, vup, vfov, aspect, aperture, focus_dist=None, t0=0.0, t1=0.0) -> Camera:
    """The product's Camera constructor (host/common-model.cpp; reference common-model.cpp:136-154)."""
    out = Camera()
    host().rtwh_camera((C.c_double * 3)(*lookfrom), (C.c_double * 3)(*lookat), (C.c_double * 3)(*vup), vfov, aspect,
                       aperture, -1.0 if focus_dist is None else float(focus_dist), t0, t1, C.byref(out))
    return out


def _from_handle(h, params, H=None) -> Scene:
    H = H or host()
    if not h:
        raise RtwError(H.rtwh_last_error().decode())
    try:
        prims = np.zeros(H.rtwh_scene_nprims(h), dtype=PRIM_DTYPE)
        mats = np.zeros(H.rtwh_scene_nmats(h), dtype=MAT_DTYPE)
        d = SceneDesc()
        H.rtwh_scene_flatten(h, prims.ctypes.data_as(_VP), mats.ctypes.data_as(_VP), C.byref(d))
        cam = Camera.from_buffer_copy(d.camera)
    finally:
        H.rtwh_scene_free(h)
    return Scene(prims, mats, cam, params)


def cover_scene(nsqrt: int = 11, aspect: float = 1.5, moving: bool = True, host_seed: int | None = 5489, H=None) -> Scene:
    """rtweekend::lots_of_balls (host/scenes.cpp; reference main.cpp:23-83).  host_seed=5489 is the state a fresh
    reference process starts from.  H: the host library to build with (default: the virtual primitive model)."""
    H = H or host()
    if host_seed is not None:
        H.rtwh_seed(host_seed)
    params = dict(lookfrom=(13, 2, 3), lookat=(0, 0, 0), vup=(0, 1, 0), vfov=20.0, aspect=aspect, aperture=0.1,
                  focus_dist=10.0, t0=0.0, t1=1.0)
    return _from_handle(H.rtwh_scene_cover(nsqrt, aspect, int(moving)), params, H)


def obj_scene(path: str, aspect: float = 1.5, H=None) -> Scene:
    """rtweekend::foo (host/scenes.cpp; reference main.cpp:85-136)."""
    H = H or host()
    params = dict(lookfrom=(1, 0, -1), lookat=(0, 0, 0), vup=(0, 1, 0), vfov=35.0, aspect=aspect, aperture=0.01,
                  focus_dist=-1.0, t0=0.0, t1=1.0)
    return _from_handle(H.rtwh_scene_obj(str(path).encode(), aspect), params, H)


def mesh_on_ground_scene(path: str, aspect: float = 1.5) -> Scene:
    s = _from_handle(host().rtwh_scene_mesh_on_ground(str(path).encode(), aspect), None)
    tri = s.prims[s.prims["kind"] == RTW_TRIANGLE]
    ys = np.concatenate([tri["a"][:, 1], tri["b"][:, 1], tri["c"][:, 1]])
    s.params = dict(lookfrom=(2.6, 1.7, 4.2), lookat=(0, 0.5 * float(ys.max() - ys.min()), 0), vup=(0, 1, 0), vfov=30.0,
                    aspect=aspect, aperture=0.02, focus_dist=-1.0, t0=0.0, t1=1.0)
    return s


def custom_scene(prims: np.ndarray, mats: np.ndarray, **cam) -> Scene:
    camera = make_camera(cam["lookfrom"], cam["lookat"], cam["vup"], cam["vfov"], cam["aspect"], cam["aperture"],
                         None if cam.get("focus_dist", -1.0) is None or cam.get("focus_dist", -1.0) <= 0 else cam["focus_dist"],
                         cam.get("t0", 0.0), cam.get("t1", 0.0))
    params = dict(cam)
    if params.get("focus_dist") is None:
        params["focus_dist"] = -1.0
    return Scene(np.asarray(prims, dtype=PRIM_DTYPE), np.asarray(mats, dtype=MAT_DTYPE), camera, params)


def image_height(width: int, aspect: float) -> int:
    return host().rtwh_image_height(width, aspect)


# ----------------------------------------------------------------------------------------------------------
# Rendering through the C ABI
# ----------------------------------------------------------------------------------------------------------
def make_cfg(width, height, spp, max_child_rays=20, sample_begin=0, kernel=KERNEL_AUTO, seed=0, device=0, stats=False,
             rays_per_lane=0, row_tiles=None, flags=0) -> RenderCfg:
    """row_tiles = (rows per tile, number of participants, index of this participant) selects the row-tile split."""
    c = RenderCfg()
    c.width, c.height = width, height
    c.sample_begin, c.sample_end = sample_begin, sample_begin + spp
    c.max_child_rays, c.kernel, c.seed, c.device = max_child_rays, kernel, seed, device
    c.flags = (FLAG_STATS if stats else 0) | flags
    if row_tiles is not None:
        c.row_tile_rows, c.row_tile_count, c.row_tile_index = row_tiles
    c.rays_per_lane = rays_per_lane
    return c


def render(scene: Scene, width: int, height: int, spp: int, max_child_rays: int = 20, **kw):
    """rtw_render: host buffers in, host accumulation buffer out.  Returns (accum[H,W,4] float32, stats dict)."""
    cfg = make_cfg(width, height, spp, max_child_rays, **kw)
    d = scene.desc()
    out = np.zeros((height, width, 4), np.float32)
    st = Stats()
    _check(lib().rtw_render(C.byref(d), C.byref(cfg), out.ctypes.data_as(_VP), C.byref(st)), "rtw_render")
    return out, st.as_dict()


def render_rgb8(scene: Scene, width: int, height: int, spp: int, max_child_rays: int = 20, ngpus: int = 1, **kw):
    """rtw_render_rgb8 / rtw_render_multi_gpu_rgb8: the quantised picture (write_color on the device).  Returns (rgb[H,W,3] uint8, stats)."""
    cfg = make_cfg(width, height, spp, max_child_rays, **kw)
    d = scene.desc()
    out = np.zeros((height, width, 3), np.uint8)
    st = Stats()
    if ngpus > 1:
        _check(lib().rtw_render_multi_gpu_rgb8(C.byref(d), C.byref(cfg), ngpus, out.ctypes.data_as(_VP), C.byref(st)), "rtw_render_multi_gpu_rgb8")
    else:
        _check(lib().rtw_render_rgb8(C.byref(d), C.byref(cfg), out.ctypes.data_as(_VP), C.byref(st)), "rtw_render_rgb8")
    return out, st.as_dict()


def scene_hash(scene: Scene) -> int:
    d = scene.desc()
    h = C.c_uint64(0)
    _check(lib().rtw_scene_hash(C.byref(d), C.byref(h)), "rtw_scene_hash")
    return h.value


def render_multi_gpu(scene: Scene, width: int, height: int, spp: int, ngpus: int, max_child_rays: int = 20, **kw):
    cfg = make_cfg(width, height, spp, max_child_rays, **kw)
    d = scene.desc()
    out = np.zeros((height, width, 4), np.float32)
    st = Stats()
    _check(lib().rtw_render_multi_gpu(C.byref(d), C.byref(cfg), ngpus, out.ctypes.data_as(_VP), C.byref(st)),
           "rtw_render_multi_gpu")
    return out, st.as_dict()


def primary_hits(scene: Scene, width: int, height: int, time: float = 0.0, precision: int = 32,
                 kernel: int = KERNEL_AUTO, device: int = 0):
    d = scene.desc()
    n = width * height
    pid = np.zeros(n, np.int32)
    t = np.zeros(n, np.float64)
    nrm = np.zeros((n, 3), np.float64)
    front = np.zeros(n, np.uint8)
    _check(lib().rtw_primary_hits(C.byref(d), width, height, time, precision, kernel, device, pid.ctypes.data_as(_VP),
                                  t.ctypes.data_as(_VP), nrm.ctypes.data_as(_VP), front.ctypes.data_as(_VP)),
           "rtw_primary_hits")
    return pid.reshape(height, width), t.reshape(height, width), nrm.reshape(height, width, 3), front.reshape(height, width)


def finalize_rgb8(accum: np.ndarray, spp: int, device: int = 0) -> np.ndarray:
    accum = np.ascontiguousarray(accum, np.float32)
    h, w = accum.shape[:2]
    out = np.zeros((h, w, 3), np.uint8)
    _check(lib().rtw_finalize_rgb8(accum.ctypes.data_as(_VP), h * w, spp, device, out.ctypes.data_as(_VP)),
           "rtw_finalize_rgb8")
    return out


def debug_scatter(mats: np.ndarray, dir_in, normal, front, ball, coin, device: int = 0):
    mats = np.ascontiguousarray(mats, dtype=MAT_DTYPE)
    n = len(mats)
    f32 = lambda a, k: np.ascontiguousarray(np.asarray(a, np.float32).reshape(n, k) if k > 1 else np.asarray(a, np.float32).reshape(n))
    dir_in, normal, ball, coin = f32(dir_in, 3), f32(normal, 3), f32(ball, 3), f32(coin, 1)
    front = np.ascontiguousarray(front, np.uint8)
    out_dir = np.zeros((n, 3), np.float32)
    out_att = np.zeros((n, 3), np.float32)
    sc = np.zeros(n, np.uint8)
    _check(lib().rtw_debug_scatter(device, n, mats.ctypes.data_as(C.POINTER(Material)), dir_in.ctypes.data_as(_VP),
                                   normal.ctypes.data_as(_VP), front.ctypes.data_as(_VP), ball.ctypes.data_as(_VP),
                                   coin.ctypes.data_as(_VP), out_dir.ctypes.data_as(_VP), out_att.ctypes.data_as(_VP),
                                   sc.ctypes.data_as(_VP)), "rtw_debug_scatter")
    return out_dir, out_att, sc


def debug_samples(n: int, seed: int = 0, device: int = 0):
    ball = np.zeros((n, 3), np.float32)
    disk = np.zeros((n, 2), np.float32)
    u = np.zeros((n, 4), np.float32)
    _check(lib().rtw_debug_samples(device, n, seed, ball.ctypes.data_as(_VP), disk.ctypes.data_as(_VP),
                                   u.ctypes.data_as(_VP)), "rtw_debug_samples")
    return ball, disk, u


def flatten_info(scene: "Scene") -> dict:
    """rtw_flatten_info: the host-side flatten + BVH build of rtw_scene_upload, without touching a GPU."""
    d = scene.desc()
    r = FlattenReport()
    _check(lib().rtw_flatten_info(C.byref(d), C.byref(r)), "rtw_flatten_info")
    return r.as_dict()


def fp32_peak(device: int = 0, seconds: float = 1.0):
    t, m = C.c_double(0), C.c_double(0)
    _check(lib().rtw_fp32_peak(device, seconds, C.byref(t), C.byref(m)), "rtw_fp32_peak")
    return t.value, m.value


class DeviceScene:
    """rtw_scene_upload / rtw_render_device: scene resident in HBM, accumulation into a caller-owned device buffer
    (a torch int64 tensor [H, W, 4] on the same device)."""

    def __init__(self, scene: Scene, device: int = 0, gpu_build: bool = False):
        """gpu_build: linear BVH built on the device (rtw_scene_upload_ex + RTW_FLAG_BVH_BUILD_GPU) instead of the host SAH tree."""
        self.scene = scene
        self.device = device
        self._h = _VP()
        d = scene.desc()
        _check(lib().rtw_scene_upload_ex(C.byref(d), device, FLAG_BVH_BUILD_GPU if gpu_build else 0, C.byref(self._h)), "rtw_scene_upload_ex")

    def check(self) -> dict:
        """rtw_scene_check: structural check of the device-resident binary tree (reads the arena back)."""
        r = FlattenReport()
        _check(lib().rtw_scene_check(self._h, C.byref(r)), "rtw_scene_check")
        return r.as_dict()

    def render_into(self, accum_fx, width, height, spp, max_child_rays=20, stream_ptr=0, want_stats=False, **kw):
        cfg = make_cfg(width, height, spp, max_child_rays, device=self.device, **kw)
        st = Stats()
        ptr = accum_fx.data_ptr() if hasattr(accum_fx, "data_ptr") else int(accum_fx)
        _check(lib().rtw_render_device(self._h, C.byref(cfg), _VP(ptr), _VP(stream_ptr),
                                       C.byref(st) if want_stats else None), "rtw_render_device")
        return st.as_dict() if want_stats else None

    def update(self, scene: Scene):
        """rtw_scene_update: new host arrays -> the same device allocation (flatten + BVH build + one H2D copy)."""
        self.scene = scene
        d = scene.desc()
        _check(lib().rtw_scene_update(self._h, C.byref(d)), "rtw_scene_update")

    def finalize_rgb8(self, accum_fx, rgb8, npixels, spp, stream_ptr=0):
        """rtw_finalize_rgb8_device: int64 sums -> uint8 rgb (both device tensors)."""
        _check(lib().rtw_finalize_rgb8_device(_VP(accum_fx.data_ptr()), npixels, spp, self.device, _VP(stream_ptr), _VP(rgb8.data_ptr())),
               "rtw_finalize_rgb8_device")

    def accum_to_float(self, accum_fx, out_f32, npixels, stream_ptr=0):
        _check(lib().rtw_accum_to_float(_VP(accum_fx.data_ptr()), _VP(out_f32.data_ptr()), npixels, self.device,
                                        _VP(stream_ptr)), "rtw_accum_to_float")

    def untile(self, gathered, accum_fx, width, height, tile_rows, count, stream_ptr=0):
        """[count][local_rows][width][4] int64 gathered buffers -> [height][width][4] (device tensors)."""
        _check(lib().rtw_untile_accum(_VP(gathered.data_ptr()), _VP(accum_fx.data_ptr()), width, height, tile_rows, count, self.device,
                                      _VP(stream_ptr)), "rtw_untile_accum")

    def close(self):
        if self._h:
            lib().rtw_scene_free(self._h)
            self._h = _VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def read_ppm(path_or_text) -> np.ndarray:
    """P3 text -> uint8 [H, W, 3]."""
    text = Path(path_or_text).read_text() if isinstance(path_or_text, (str, os.PathLike)) and "\n" not in str(path_or_text) else path_or_text
    tok = text.split()
    assert tok[0] == "P3" and tok[3] == "255"
    w, h = int(tok[1]), int(tok[2])
    return np.array(tok[4:4 + 3 * w * h], dtype=np.int64).astype(np.uint8).reshape(h, w, 3)


def quantize(accum: np.ndarray, spp: int) -> np.ndarray:
    """write_color (reference render.cpp:11-20) in numpy double: for comparing images, not a render path."""
    c = np.sqrt(np.asarray(accum, np.float64)[..., :3] / float(spp))
    return (256 * np.clip(c, 0.0, 0.999)).astype(np.int64).astype(np.uint8)


# ----------------------------------------------------------------------------------------------------------
# Multi-GPU plumbing (one process per GPU, torch.distributed): samples-per-pixel sharded over ranks, ONE reduce.
# Mirrors the reference's thread fan-out + image sum (render.cpp:169-180).
# ----------------------------------------------------------------------------------------------------------
FIXED_POINT_ONE = float(1 << 32)


def sample_shard(spp: int, rank: int, world: int) -> tuple[int, int]:
    """Global sample indices [begin, end) of `rank`.  Requires world | spp: the reference analogue silently drops the
    remainder (spp / nthreads, render.cpp:174); here that is an error."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    if spp % world != 0:
        raise ValueError(f"samples_per_pixel={spp} does not split evenly over {world} GPUs")
    per = spp // world
    return rank * per, (rank + 1) * per


def row_tile_local_rows(height: int, tile_rows: int, count: int) -> int:
    """Rows of one participant's packed buffer in a row-tile split (same formula as rtw_row_tile_local_rows)."""
    if height < 1 or tile_rows < 1 or count < 1:
        raise ValueError("bad row-tile split")
    tiles = -(-height // tile_rows)
    return -(-tiles // count) * tile_rows


def untile_rows(gathered, height: int, tile_rows: int, count: int):
    """Host/torch restatement of rtw_untile_accum for tensors on any device: gathered [count, local_rows, W, C] ->
    [height, W, C].  Tile t of the image belongs to participant t % count and is its (t // count)-th tile."""
    local_rows = gathered.shape[1]
    tiles_local = local_rows // tile_rows
    g = gathered.reshape(count, tiles_local, tile_rows, *gathered.shape[2:])
    full = g.transpose(0, 1).reshape(tiles_local * count * tile_rows, *gathered.shape[2:])
    return full[:height]


def gather_row_tiles(accum_local, dst: int = 0):
    """The collective of the row-tile alternative: the packed per-rank buffers are gathered on `dst`, no summation.
    Returns the [world, local_rows, W, 4] tensor on dst (None elsewhere); world 1 is a view of the input."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return accum_local.unsqueeze(0)
    world, rank = dist.get_world_size(), dist.get_rank()
    out = torch.empty((world, *accum_local.shape), dtype=accum_local.dtype, device=accum_local.device) if rank == dst else None
    dist.gather(accum_local, list(out.unbind(0)) if rank == dst else None, dst=dst)
    return out


def reduce_accum(accum_fx, dst: int = 0):
    """The one collective of the path: integer sum of the int64 fixed-point accumulation buffers onto `dst`
    (NCCL over NVLink on GPUs, gloo on CPU in the tests).  Exact, hence independent of world size and order."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum_fx, dst=dst, op=dist.ReduceOp.SUM)
    return accum_fx
