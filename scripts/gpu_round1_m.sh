#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
# memory checker on small configurations of every kernel; one sanitizer tool per call
cat > /tmp/san.py <<'PY'
import importlib, sys
sys.path.insert(0, '.')
rtw = importlib.import_module('raytracing-one-weekend_b200')
sc = rtw.cover_scene()
for k, r in ((rtw.KERNEL_SPHERES_SMEM, 1), (rtw.KERNEL_SPHERES_SMEM, 2), (rtw.KERNEL_SPHERES_SMEM, 4), (rtw.KERNEL_BVH, 0), (rtw.KERNEL_BVH, 1000)):
    acc, st = rtw.render(sc, 64, 42, 4, 20, kernel=k, rays_per_lane=r, stats=True)
    assert st['paths'] == 64 * 42 * 4
    rtw.primary_hits(sc, 64, 42, 0.5, 32, k)
rtw.primary_hits(sc, 64, 42, 0.5, 64)
ms = rtw.mesh_on_ground_scene('tests/golden/suzanne.obj')
rtw.render(ms, 64, 42, 4, 20)
rtw.primary_hits(ms, 64, 42, 0.0, 32)
rtw.finalize_rgb8(acc, 4)
rtw.debug_samples(1000)
print('sanitizer workload done')
PY
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
tail -6 gpurun_out/sanitizer_memcheck.log
