import importlib, sys, time, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
rtw = importlib.import_module("raytracing-one-weekend_b200")
n = C.c_longlong(0)
rtw.host().rtwh_make_mesh(b'assets/suzanne.obj', b'/tmp/standin5.obj', 5, 20221018, 0.08, C.byref(n))
t0 = time.perf_counter(); scene = rtw.mesh_on_ground_scene('/tmp/standin5.obj', 1.7777777777777777); print(f"scene build (OBJ parse + host model + flatten) {time.perf_counter()-t0:.2f} s, {len(scene.prims)} prims")
import numpy as np
out = np.zeros((1080, 1920, 4), np.float32)
for spp in (4, 4, 64):
    cfg = rtw.make_cfg(1920, 1080, spp, 20)
    t0 = time.perf_counter(); d = scene.desc(); t1 = time.perf_counter()
    st = rtw.Stats()
    rc = rtw.lib().rtw_render(C.byref(d), C.byref(cfg), out.ctypes.data_as(C.c_void_p), C.byref(st)); t2 = time.perf_counter()
    print(f"spp {spp}: desc() {1e3*(t1-t0):.1f} ms, call wall {1e3*(t2-t1):.1f} ms, inside: total {st.total_ms:.1f}, upload {st.h2d_ms:.1f}, kernel {st.kernel_ms:.1f}, d2h {st.d2h_ms:.1f}")
