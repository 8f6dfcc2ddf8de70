"""TEST INFRASTRUCTURE ONLY.  ctypes bindings for the two CPU oracles:

  * ``port``  -- oracle/librtw_oracle.so, the plain-C restatement (oracle/rtw_oracle.c), built anywhere with gcc;
  * ``ref``   -- oracle/_ref/libref_oracle.so, the UNMODIFIED reference sources compiled against header shims
                 (oracle/ref_harness.cpp + oracle/Makefile), built only where /root/reference exists and shipped
                 as a prebuilt file to the GPU box.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The product (raytracing-one-weekend_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
PORT_LIB = HERE / "librtw_oracle.so"
REF_LIB = HERE / "_ref" / "libref_oracle.so"
REF_EXE = HERE / "_ref" / "rtweekend_ref"

PRIM_DTYPE = np.dtype([("kind", "<i4"), ("material", "<i4"), ("a", "<f8", 3), ("b", "<f8", 3), ("c", "<f8", 3),
                       ("radius", "<f8")])
MAT_DTYPE = np.dtype([("kind", "<i4"), ("pad", "<i4"), ("albedo", "<f8", 3), ("fuzz", "<f8"), ("ior", "<f8")])


class CameraParams(C.Structure):
    _fields_ = [("lookfrom", C.c_double * 3), ("lookat", C.c_double * 3), ("vup", C.c_double * 3),
                ("vfov", C.c_double), ("aspect", C.c_double), ("aperture", C.c_double), ("focus_dist", C.c_double),
                ("t0", C.c_double), ("t1", C.c_double)]


def camera_params(lookfrom, lookat, vup, vfov, aspect, aperture, focus_dist=-1.0, t0=0.0, t1=0.0) -> CameraParams:
    p = CameraParams()
    p.lookfrom[:] = lookfrom
    p.lookat[:] = lookat
    p.vup[:] = vup
    p.vfov, p.aspect, p.aperture = vfov, aspect, aperture
    p.focus_dist = -1.0 if focus_dist is None else focus_dist
    p.t0, p.t1 = t0, t1
    return p


def build(ref: bool = True) -> None:
    """make -C oracle port [ref]: building the checker is not using it."""
    targets = ["port"] + (["ref"] if ref and Path("/root/reference/src/render.cpp").exists() else [])
    r = subprocess.run(["make", "-C", str(HERE), *targets], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"oracle build failed:\n{r.stdout}\n{r.stderr}")


_VP = C.c_void_p


class _Oracle:
    """Common surface of both oracles (function-name prefix differs)."""

    def __init__(self, path: Path, prefix: str):
        if not path.exists():
            raise FileNotFoundError(f"{path} not built (run oracle.build())")
        self.L = C.CDLL(str(path))
        self.p = prefix
        f = self._f
        f("scene_cover").restype = _VP
        f("scene_cover").argtypes = [C.c_int, C.c_double, C.c_int]
        f("scene_obj").restype = _VP
        f("scene_obj").argtypes = [C.c_char_p, C.c_double]
        f("scene_custom").restype = _VP
        f("scene_custom").argtypes = [_VP, C.c_int, _VP, C.c_int, C.POINTER(CameraParams)]
        f("scene_free").argtypes = [_VP]
        f("scene_nprims").argtypes = [_VP]
        f("scene_nmats").argtypes = [_VP]
        f("scene_dump").argtypes = [_VP, _VP, _VP]
        f("scene_camera").argtypes = [_VP, C.POINTER(CameraParams)]
        f("seed").argtypes = [C.c_uint32]
        f("random_double").restype = C.c_double

    def _f(self, name):
        return getattr(self.L, self.p + name)

    # -- RNG of the reference host ------------------------------------------------------------------------
    def seed(self, s: int = 5489):
        self._f("seed")(s)

    def random_double(self) -> float:
        return self._f("random_double")()

    # -- scenes ---------------------------------------------------------------------------------------------
    def scene_cover(self, nsqrt=11, aspect=1.5, moving=True, seed=5489):
        if seed is not None:
            self.seed(seed)
        return OracleScene(self, self._f("scene_cover")(nsqrt, aspect, int(moving)))

    def scene_obj(self, path, aspect=1.5, seed=5489):
        if seed is not None:
            self.seed(seed)
        h = self._f("scene_obj")(str(path).encode(), aspect)
        if not h:
            raise RuntimeError(f"oracle could not load {path}")
        return OracleScene(self, h)

    def scene_custom(self, prims, mats, cam: CameraParams):
        prims = np.ascontiguousarray(prims, dtype=PRIM_DTYPE)
        mats = np.ascontiguousarray(mats.view(MAT_DTYPE) if mats.dtype != MAT_DTYPE else mats)
        return OracleScene(self, self._f("scene_custom")(prims.ctypes.data_as(_VP), len(prims), mats.ctypes.data_as(_VP),
                                                         len(mats), C.byref(cam)))


class OracleScene:
    def __init__(self, o: _Oracle, handle):
        self.o, self.h = o, _VP(handle)

    def close(self):
        if self.h:
            self.o._f("scene_free")(self.h)
            self.h = _VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def nprims(self):
        return self.o._f("scene_nprims")(self.h)

    @property
    def nmats(self):
        return self.o._f("scene_nmats")(self.h)

    def dump(self):
        prims = np.zeros(self.nprims, PRIM_DTYPE)
        mats = np.zeros(self.nmats, MAT_DTYPE)
        self.o._f("scene_dump")(self.h, prims.ctypes.data_as(_VP), mats.ctypes.data_as(_VP))
        return prims, mats

    def camera(self) -> CameraParams:
        c = CameraParams()
        self.o._f("scene_camera")(self.h, C.byref(c))
        return c

    def primary_hits(self, width, height, time=0.0, nthreads=1):
        n = width * height
        pid = np.zeros(n, np.int32)
        t = np.zeros(n)
        nrm = np.zeros((n, 3))
        front = np.zeros(n, np.uint8)
        if nthreads > 1 and self.o.p == "rtwo_":   # the port can split the rows over threads (big meshes, brute force)
            fn = self.o._f("primary_hits_mt")
            fn.argtypes = [_VP, C.c_int, C.c_int, C.c_double, C.c_int, _VP, _VP, _VP, _VP]
            fn.restype = None
            fn(self.h, width, height, time, nthreads, pid.ctypes.data_as(_VP), t.ctypes.data_as(_VP), nrm.ctypes.data_as(_VP),
               front.ctypes.data_as(_VP))
            self.bvh_vs_bruteforce_disagreements = 0
            return pid.reshape(height, width), t.reshape(height, width), nrm.reshape(height, width, 3), front.reshape(height, width)
        fn = self.o._f("primary_hits")
        fn.argtypes = [_VP, C.c_int, C.c_int, C.c_double, _VP, _VP, _VP, _VP]
        fn.restype = C.c_int
        rc = fn(self.h, width, height, time, pid.ctypes.data_as(_VP), t.ctypes.data_as(_VP), nrm.ctypes.data_as(_VP),
                front.ctypes.data_as(_VP))
        self.bvh_vs_bruteforce_disagreements = rc if self.o.p == "ref_" else 0
        return pid.reshape(height, width), t.reshape(height, width), nrm.reshape(height, width, 3), front.reshape(height, width)

    def render_linear(self, width, height, spp, max_child_rays=20, seed=None, want_sumsq=True):
        """Reference sampling order on the global mt19937, one thread: (sum[H,W,3], sumsq[H,W,3], rays)."""
        if seed is not None:
            self.o.seed(seed)
        s = np.zeros((height, width, 3))
        q = np.zeros((height, width, 3)) if want_sumsq else None
        fn = self.o._f("render_linear")
        if self.o.p == "ref_":
            fn.argtypes = [_VP, C.c_int, C.c_int, C.c_int, C.c_int, _VP, _VP]
            fn(self.h, width, height, spp, max_child_rays, s.ctypes.data_as(_VP), q.ctypes.data_as(_VP))
            return s, q, None
        rays = C.c_uint64(0)
        fn.argtypes = [_VP, C.c_int, C.c_int, C.c_int, C.c_int, _VP, _VP, C.POINTER(C.c_uint64)]
        fn(self.h, width, height, spp, max_child_rays, s.ctypes.data_as(_VP), q.ctypes.data_as(_VP) if q is not None else None,
           C.byref(rays))
        return s, q, rays.value


class Port(_Oracle):
    def __init__(self):
        super().__init__(PORT_LIB, "rtwo_")
        self.L.rtwo_random_double_range.restype = C.c_double
        self.L.rtwo_random_double_range.argtypes = [C.c_double, C.c_double]

    def render_philox(self, scene: OracleScene, width, height, sample_begin, sample_end, max_child_rays=20, seed=0,
                      nthreads=1, want_sumsq=False):
        s = np.zeros((height, width, 3))
        q = np.zeros((height, width, 3)) if want_sumsq else None
        rays = C.c_uint64(0)
        fn = self.L.rtwo_render_philox
        fn.argtypes = [_VP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, _VP, _VP, C.POINTER(C.c_uint64)]
        fn(scene.h, width, height, sample_begin, sample_end, max_child_rays, seed, nthreads, s.ctypes.data_as(_VP),
           q.ctypes.data_as(_VP) if q is not None else None, C.byref(rays))
        return s, q, rays.value

    def quantize(self, sums, spp):
        sums = np.ascontiguousarray(sums, np.float64)
        h, w = sums.shape[:2]
        out = np.zeros((h, w, 3), np.uint8)
        self.L.rtwo_quantize.argtypes = [_VP, C.c_int, C.c_int, _VP]
        self.L.rtwo_quantize(sums.ctypes.data_as(_VP), h * w, spp, out.ctypes.data_as(_VP))
        return out

    def philox(self, ctr, key):
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        self.L.rtwo_philox4x32_10(c, k, o)
        return list(o)

    def hit_sphere(self, org, d, tmin, tmax, center, radius):
        t = C.c_double(0)
        p = (C.c_double * 3)()
        n = (C.c_double * 3)()
        fr = C.c_int(0)
        self.L.rtwo_hit_sphere.argtypes = [C.c_double * 3, C.c_double * 3, C.c_double, C.c_double, C.c_double * 3, C.c_double,
                                           C.POINTER(C.c_double), C.c_double * 3, C.c_double * 3, C.POINTER(C.c_int)]
        ok = self.L.rtwo_hit_sphere((C.c_double * 3)(*org), (C.c_double * 3)(*d), tmin, tmax, (C.c_double * 3)(*center), radius,
                                    C.byref(t), p, n, C.byref(fr))
        return (t.value, np.array(p), np.array(n), bool(fr.value)) if ok else None

    def hit_triangle(self, org, d, tmin, tmax, a, b, c):
        t = C.c_double(0)
        p = (C.c_double * 3)()
        n = (C.c_double * 3)()
        v = lambda x: (C.c_double * 3)(*x)
        self.L.rtwo_hit_triangle.argtypes = [C.c_double * 3] * 2 + [C.c_double, C.c_double] + [C.c_double * 3] * 3 + \
            [C.POINTER(C.c_double), C.c_double * 3, C.c_double * 3]
        ok = self.L.rtwo_hit_triangle(v(org), v(d), tmin, tmax, v(a), v(b), v(c), C.byref(t), p, n)
        return (t.value, np.array(p), np.array(n)) if ok else None

    def scatter(self, mat_record, dir_in, normal, front, ball, coin):
        m = np.zeros(1, MAT_DTYPE)
        m[0] = mat_record
        d = (C.c_double * 3)()
        a = (C.c_double * 3)()
        v = lambda x: (C.c_double * 3)(*[float(y) for y in x])
        self.L.rtwo_scatter.argtypes = [_VP, C.c_double * 3, C.c_double * 3, C.c_int, C.c_double * 3, C.c_double, C.c_double * 3,
                                        C.c_double * 3]
        ok = self.L.rtwo_scatter(m.ctypes.data_as(_VP), v(dir_in), v(normal), int(front), v(ball), float(coin), d, a)
        return (np.array(d), np.array(a)) if ok else None

    def sky(self, d):
        out = (C.c_double * 3)()
        self.L.rtwo_sky.argtypes = [C.c_double * 3, C.c_double * 3]
        self.L.rtwo_sky((C.c_double * 3)(*d), out)
        return np.array(out)


class Ref(_Oracle):
    def __init__(self):
        super().__init__(REF_LIB, "ref_")


_port = None
_ref = None


def port() -> Port:
    global _port
    if _port is None:
        if not PORT_LIB.exists():
            build(ref=False)
        _port = Port()
    return _port


def ref() -> Ref:
    global _ref
    if _ref is None:
        _ref = Ref()
    return _ref


def ref_available() -> bool:
    return REF_LIB.exists() and REF_EXE.exists()
