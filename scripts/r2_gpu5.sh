#!/bin/bash
# GPU-built BVH: validity + parity test, build time and trace speed against the host SAH tree on the 991k-triangle mesh; cold-start trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "built_on_the_device or cache or stand_in" 2>&1 | tail -5
python - <<'PY'
import importlib, ctypes as C, sys, time
sys.path.insert(0,'.')
rtw=importlib.import_module('raytracing-one-weekend_b200')
n=C.c_longlong(0); rtw.host().rtwh_make_mesh(b'assets/suzanne.obj', b'/tmp/standin5.obj', 5, 20221018, 0.08, C.byref(n))
aspect=1.7777777777777777
for name, scene in (("standin_991k", rtw.mesh_on_ground_scene('/tmp/standin5.obj', aspect)), ("suzanne", rtw.mesh_on_ground_scene('assets/suzanne.obj', aspect)), ("grid_57k", rtw.cover_scene(120, aspect))):
    import os
    for label, fl, env in (("host SAH", rtw.FLAG_BVH_BUILD_HOST, "1"), ("device LBVH", rtw.FLAG_BVH_BUILD_GPU, "0"), ("LBVH+SAH top", rtw.FLAG_BVH_BUILD_GPU, "1")):
        os.environ["RTW_LBVH_SAH_TOP"] = env
        best=None
        for it in range(3):
            t=time.perf_counter(); acc, st = rtw.render(scene, 1920, 1080, 64, 20, seed=1, flags=fl | rtw.FLAG_NO_SCENE_CACHE); dt=(time.perf_counter()-t)*1e3
            best = (dt, st) if best is None or dt < best[0] else best
        dt, st = best
        print(f"{name:14s} {label:12s} call {dt:7.1f} ms: scene {st['h2d_ms']:6.1f} ms (device build {st['bvh_build_gpu_ms']:.2f}), kernel {st['kernel_ms']:6.2f} ms = {st['paths']/st['kernel_ms']/1e3:7.1f} Mpaths/s, d2h {st['d2h_ms']:.1f}")
PY
EXE=./raytracing-one-weekend_b200/rtweekend
RTW_TRACE=1 $EXE -w 1920 -a 1.7777777777777777 -s 256 -l /tmp/standin5.obj --scene mesh-on-ground > /dev/null 2> gpurun_out/trace_mesh.err; grep -E "trace|host:|Done|kernel" gpurun_out/trace_mesh.err | tr '\n' ' '; echo
timeout 900 python bench.py --steps 3 --warmup 3 --workload dragon_standin_1080p_256spp > gpurun_out/bench_dragon.log 2>&1; tail -1 gpurun_out/bench_dragon.log | cut -c1-2500
