#!/usr/bin/env python
"""bench.py -- Mpaths/s of the per-pixel / per-sample path-tracing loop on BASELINE.json's config 2
(cover scene, 1920x1080, 1024 spp, max depth 50), 1..8 B200, next to the reference CPU renderer.

    python bench.py [--gpus N] [--steps K] [--warmup W]             # our arm (torchrun launches one rank per GPU)
    python bench.py --impl reference [--gpus N] [--steps K] ...     # the reference's own CPU implementation

A step = one full render of the workload: zero the accumulation buffer, trace every (pixel, sample) path of this
rank's sample shard, combine the shards with one NCCL reduce (N > 1), convert to the float accumulation buffer on
GPU 0.  `value` = paths of the whole job / device time (inputs resident in HBM); `e2e` = the same through the host-buffer
C-ABI call with host<->device copies inside the timed region.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (width, aspect, spp, max_child_rays, nsqrt, moving)
    "cover_1080p_1024spp_depth50": (1920, 1.7777777777777777, 1024, 50, 11, True),   # BASELINE.json configs[1]
    "cover_default_200x133_20spp_depth20": (200, 1.5, 20, 20, 11, True),            # configs[0]
    "cover_4k_4096spp_depth50": (3840, 1.7777777777777777, 4096, 50, 11, True),     # configs[4]
    # mesh workloads: (width, aspect, spp, max_child_rays, subdivision rounds of suzanne.obj, -)
    "suzanne_on_ground_1080p_256spp": (1920, 1.7777777777777777, 256, 20, 0, None),  # configs[2]
    "dragon_standin_1080p_256spp": (1920, 1.7777777777777777, 256, 20, 5, None),     # configs[3]: dragon.obj is missing from the
}                                                                                    # reference mount; 968*4^5 = 991,232-triangle stand-in
SUZANNE = ROOT / "assets" / "suzanne.obj"


def mesh_path(rtw, wl):
    """OBJ file of a mesh workload: suzanne.obj itself, or the stand-in generated from it (cached in the temp directory)."""
    a = wl[4]
    if a <= 0:
        return str(SUZANNE)
    import ctypes as C
    import tempfile
    path = os.path.join(tempfile.gettempdir(), f"rtw_standin_r{a}.obj")
    if not os.path.exists(path):
        n = C.c_longlong(0)
        if rtw.host().rtwh_make_mesh(str(SUZANNE).encode(), path.encode(), a, 20221018, 0.08, C.byref(n)) != 0:
            raise RuntimeError(rtw.host().rtwh_last_error().decode())
    return path


def build_scene(rtw, name, wl):
    """The product's own scene builders (host C++): cover scene, or a mesh standing on the r=1000 ground sphere."""
    width, aspect, spp, depth, a, b = wl
    if name.startswith("cover"):
        return rtw.cover_scene(a, aspect, b)
    return rtw.mesh_on_ground_scene(mesh_path(rtw, wl), aspect)
# Per-launch figures of the dominant kernel come from the `ncu --set full` summaries committed under profiles/ (one launch each,
# written by scripts/ncu_summary.py; newest round first), parsed here -- nothing is pasted into this file:
#   traffic       = dram__bytes_read.sum + dram__bytes_write.sum of that launch.  It is the accumulation buffer (66 MB at 1080p)
#                   being read after the memset and partly written back, plus the scene once: it does not scale with spp
#                   (profiles/r02_prof_k2w_1024spp.txt measures it at the benched size).
#   tinst_per_ray = smsp__inst_executed.sum x smsp__thread_inst_executed_per_inst_executed.ratio / rays of that launch (the
#                   `# launch:` line of the summary): the thread instructions one ray costs, numerator of the issue roofline.
PROFILE_KEYS = {("cover", "wf"): "k2w", ("cover", "perlane"): "k2_perlane", ("cover", "sweep"): "k1", ("dragon", "perlane"): "k2_dragon",
                ("suzanne", "perlane"): "k2_suzanne"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def load_profile(family: str, kernel_key: str, prefer: str = "") -> dict:
    """Parses profiles/rNN_prof_<key>[<prefer>].txt (newest round that has it).  Returns {} when there is none."""
    key = PROFILE_KEYS.get((family, kernel_key))
    if key is None:
        return {}
    cands = sorted((ROOT / "profiles").glob(f"r*_prof_{key}{prefer}.txt"), reverse=True)
    if not cands:
        return {}
    path = cands[0]
    m, launch = {}, {}
    for ln in path.read_text().splitlines():
        if ln.startswith("# launch:"):
            launch = dict(kv.split("=") for kv in ln[len("# launch:"):].split())
            continue
        f = ln.split()
        if len(f) >= 2 and not ln.startswith("#"):
            try:
                m[f[0]] = (float(f[1]), f[2] if len(f) > 2 else "")
            except ValueError:
                pass
    def val(name, default=None):
        return m[name][0] if name in m else default
    def nbytes(name):
        return m[name][0] * UNIT.get(m[name][1], 1.0) if name in m else None
    out = {"file": str(path.relative_to(ROOT))}
    rd, wr = nbytes("dram__bytes_read.sum"), nbytes("dram__bytes_write.sum")
    if rd is not None and wr is not None:
        out["traffic"] = rd + wr
    inst, lanes = val("smsp__inst_executed.sum"), val("smsp__thread_inst_executed_per_inst_executed.ratio")
    if lanes is not None:
        out["lanes_per_inst"] = lanes
    if inst is not None and lanes is not None and "rays" in launch:
        out["tinst_per_ray"] = inst * lanes / float(launch["rays"])
    # round-1 summaries carry no `# launch:` line; their ray counts are those of the frames named in profiles/README.md
    elif inst is not None and lanes is not None and path.name.startswith("r01_"):
        rays = {"k2w": 37.68e6, "k2_perlane": 37.68e6, "k1": 37.68e6, "k2_dragon": 15.0e6, "k2_suzanne": 29.7e6}.get(key)
        if rays:
            out["tinst_per_ray"] = inst * lanes / rays
    ia = val("smsp__issue_active.avg.pct_of_peak_sustained_active")
    if ia is not None:
        out["issue_active_pct"] = ia
    red = val("lts__t_sectors_op_red.sum")
    if red is not None:
        out["l2_red_sectors"] = red
    bc = val("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
    if bc is not None:
        out["smem_bank_conflicts"] = bc
    return out


# canonical FP32 flop costs of SURVEY.md 8(d) (FMA = 2)
FLOP_STATIC_TEST, FLOP_MOVING_TEST, FLOP_HIT, FLOP_SHADE = 17.0, 23.0, 40.0, 80.0
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(width: int, aspect: float, spp_per_thread: int, depth: int, threads: int) -> dict:
    """Times the reference's own CPU implementation (oracle/_ref/rtweekend_ref = its sources compiled unmodified) on a
    bounded sample of the workload; falls back to the plain-C port of the oracle when the prebuilt reference is absent."""
    import oracle
    height = int(width / aspect)
    if oracle.ref_available():
        spp = spp_per_thread * threads
        t0 = time.perf_counter()
        with open(os.devnull, "wb") as null:
            p = subprocess.run([str(oracle.REF_EXE), "-w", str(width), "-a", repr(aspect), "-s", str(spp), "-c", str(depth), "-t", str(threads)],
                               stdout=null, stderr=subprocess.PIPE)
        wall = time.perf_counter() - t0
        m = re.findall(rb"Done in (\d+)ms", p.stderr)
        if p.returncode != 0 or not m:
            raise RuntimeError("reference executable failed: " + p.stderr[-300:].decode(errors="replace"))
        secs = int(m[-1]) / 1e3  # the reference's own timer (render.cpp:188-190): BVH build + render + PPM write
        paths = width * height * spp
        return {"value": paths / secs / 1e6, "unit": "Mpaths/s", "cores": threads, "kind": "reference", "seconds": secs, "wall_seconds": wall,
                "sample": f"cover scene {width}x{height}, {spp} spp ({spp_per_thread}/thread x {threads} threads), depth {depth}; "
                          f"time = the reference's own 'Done in' timer (includes BVH build and P3 write)"}
    port = oracle.port()
    sc = port.scene_cover(11, aspect, True)
    t0 = time.perf_counter()
    _, _, rays = port.render_philox(sc, width, height, 0, spp_per_thread, depth, seed=1, nthreads=threads)
    secs = time.perf_counter() - t0
    paths = width * height * spp_per_thread
    return {"value": paths / secs / 1e6, "unit": "Mpaths/s", "cores": threads, "kind": "port", "seconds": secs,
            "sample": f"cover scene {width}x{height}, {spp_per_thread} spp, depth {depth}, oracle port (brute-force closest hit), rows split over {threads} threads"}


def cpu_reference_mesh(scene, aspect, depth) -> dict:
    """Mesh workloads: the reference has no mesh-on-ground scene builder, so its hit/scatter/BVH code (oracle/_ref, or the
    port) is driven through its public Scene API on a bounded sample, one thread (its render() would share one racy RNG)."""
    import oracle
    o = oracle.ref() if oracle.ref_available() else oracle.port()
    osc = o.scene_custom(scene.prims, scene.mats.view(oracle.MAT_DTYPE), oracle.camera_params(**scene.params))
    w, h = 160, int(160 / aspect)
    times = []
    for q in (1, 3):  # two sample counts: the difference isolates the per-path rate from the (one-off) BVH build
        t0 = time.perf_counter()
        osc.render_linear(w, h, q, depth, seed=1, want_sumsq=True)
        times.append(time.perf_counter() - t0)
    per_path = max(times[1] - times[0], 1e-9) / (w * h * 2)
    build = max(times[0] - per_path * w * h, 0.0)
    return {"value": 1.0 / per_path / 1e6, "unit": "Mpaths/s", "cores": 1, "kind": "reference" if oracle.ref_available() else "port",
            "seconds": sum(times), "bvh_build_seconds": build,
            "sample": f"same scene through the reference Scene API, {w}x{h}, 1 and 3 spp, depth {depth}, one thread; rate from the "
                      f"difference (BVH build of {build:.2f} s excluded)"}


def run_reference_arm(args, wl_name, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    width, aspect, spp, depth, _, _ = wl
    if not wl_name.startswith("cover"):
        import importlib as _il
        rtw = _il.import_module("raytracing-one-weekend_b200")
        r = cpu_reference_mesh(build_scene(rtw, wl_name, wl), aspect, depth)
        print(json.dumps({"impl": "reference", "metric": "Mpaths/s", "value": r["value"], "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": 1,
                          "warmup": 0, "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic", "config": {"workload": wl_name, "bounded_sample": r["sample"]}, "cpu_baseline": r,
                          "e2e": {"value": r["value"], "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return 0
    threads = min(os.cpu_count() or 1, 64)  # all threads share ONE unsynchronised mt19937: more only adds contention
    # bounded sample: quarter-resolution frame, 1 sample per thread (cost is linear in pixels x spp)
    bw = max(width // 2, 200)
    vals, secs = [], []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_run(bw, aspect, 1, depth, threads)
        if i >= args.warmup:
            vals.append(r["value"]); secs.append(r["seconds"])
    value = sum(v * s for v, s in zip(vals, secs)) / sum(secs)
    line = {"impl": "reference", "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": wl_name, "bounded_sample": r["sample"]},
            "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cover_1080p_1024spp_depth50", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (a reduced-size run is NOT the headline)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "spheres", "bvh", "bvh-perlane"])
    ap.add_argument("--rays-per-lane", type=int, default=0)
    ap.add_argument("--split", default="spp", choices=["spp", "rows"], help="multi-GPU work split: samples + one reduce (default), or interleaved row tiles + one gather (SURVEY 8(e) alternative)")
    ap.add_argument("--tile-rows", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cold", action="store_true", help="skip the fresh-process run of the drop-in executable")
    ap.add_argument("--traffic-bytes", type=float, default=None, help="override: dram bytes per launch of the render kernel from an ncu --set full capture")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference_arm(args, args.workload, wl)

    import numpy as np
    import torch
    import torch.distributed as dist
    rtw = importlib.import_module("raytracing-one-weekend_b200")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this renderer has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; reporting n_gpus={world}", file=sys.stderr)

    width, aspect, spp, depth, nsqrt, moving = wl
    if args.spp:
        spp = args.spp
    height = rtw.image_height(width, aspect)
    kernel = {"auto": rtw.KERNEL_AUTO, "spheres": rtw.KERNEL_SPHERES_SMEM, "bvh": rtw.KERNEL_BVH, "bvh-perlane": rtw.KERNEL_BVH_PERLANE}[args.kernel]
    scene = build_scene(rtw, args.workload, wl)
    scene_has_triangles = bool((scene.prims["kind"] == rtw.RTW_TRIANGLE).any())
    rows = args.split == "rows" and world > 1
    s_begin, s_end = (0, spp) if rows else rtw.sample_shard(spp, rank, world)
    row_tiles = (args.tile_rows, world, rank) if rows else None
    ds = rtw.DeviceScene(scene, local_rank)
    npix = width * height
    # row split: this rank's packed tiles; rank 0 also holds the assembled image
    local_rows = rtw.row_tile_local_rows(height, args.tile_rows, world) if rows else height
    accum = torch.zeros((local_rows, width, 4), dtype=torch.int64, device=dev)
    accum_full = torch.zeros((height, width, 4), dtype=torch.int64, device=dev) if rows and rank == 0 else None
    out_f32 = torch.zeros((height, width, 4), dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream()

    def step(collect_stats=False):
        accum.zero_()
        st = ds.render_into(accum, width, height, s_end - s_begin, depth, sample_begin=s_begin, stream_ptr=stream.cuda_stream, seed=0,
                            kernel=kernel, rays_per_lane=args.rays_per_lane, want_stats=collect_stats, row_tiles=row_tiles)
        combine()
        return st

    def combine():
        """The one collective of the step, then the float accumulation buffer on rank 0."""
        if rows:
            g = rtw.gather_row_tiles(accum, dst=0)
            if rank == 0:
                ds.untile(g, accum_full, width, height, args.tile_rows, world, stream_ptr=stream.cuda_stream)
                ds.accum_to_float(accum_full, out_f32, npix, stream_ptr=stream.cuda_stream)
        else:
            rtw.reduce_accum(accum, dst=0)
            if rank == 0:
                ds.accum_to_float(accum, out_f32, npix, stream_ptr=stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    # one untimed instrumented pass: exact ray count of this shard (deterministic for a given seed) and the kernel's own time
    st = step(collect_stats=True)
    rays_local = st["rays"]
    kernel_used = st["kernel_used"]
    wavefront = st.get("bvh_variant") == rtw.BVH_WAVEFRONT
    kernel_label = ("k_render_sweep<2> (K1 shared-memory sphere sweep)" if kernel_used == rtw.KERNEL_SPHERES_SMEM else
                    "k_render_wf (K2w: BVH traversal, wavefront per warp, tables + path records in shared memory)" if wavefront else
                    "k_render_bvh (K2: BVH traversal, per-lane resumable state machine" + (", tables in shared memory)" if not scene_has_triangles and len(scene.prims) < 600 else ", tables in L1/L2)"))
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    launches_before = rtw.kernel_launches()   # counted inside librtw_b200.so at every <<<>>>
    for a, k, b in ev:
        flush.fill_(1.0)  # evict L2 between timed iterations (outside the timed events)
        a.record(stream)
        accum.zero_()
        ds.render_into(accum, width, height, s_end - s_begin, depth, sample_begin=s_begin, stream_ptr=stream.cuda_stream, seed=0, kernel=kernel,
                       rays_per_lane=args.rays_per_lane, row_tiles=row_tiles)
        k.record(stream)   # end of the render kernel (for the roofline)
        combine()
        b.record(stream)
    barrier()
    launches_timed = rtw.kernel_launches() - launches_before
    clocks = sampler.stop() if rank == 0 else None
    total_ms = sum(a.elapsed_time(b) for a, _, b in ev)
    kernel_ms = sum(a.elapsed_time(k) for a, k, _ in ev) / args.steps  # zero_ (~10 us) + k_render
    t = torch.tensor([total_ms, kernel_ms, float(rays_local), float(launches_timed)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, kernel_ms, rays_total, launches_timed = tmax[0].item(), tmax[1].item(), tsum[2].item(), int(tsum[3].item())
    else:
        rays_total = float(rays_local)
    paths_total = float(npix) * spp
    ms_per_step = total_ms / args.steps
    value = paths_total / (ms_per_step * 1e-3) / 1e6

    # ---- end to end: host buffers in, host buffer out, copies inside the timed region --------------------------------------------
    # One GPU: the host-buffer C-ABI call rtw_render with RTW_FLAG_NO_SCENE_CACHE, i.e. every step flattens the host arrays, builds the
    # BVH, uploads, renders and downloads the float accumulation buffer (`e2e`); the same call with the library's scene cache
    # (`e2e_cached`: a repeated call with unchanged arrays, what progressive slices pay) and with the picture finished on the device
    # (`e2e_rgb8`: 3 bytes per pixel come back) are reported beside it.
    # N GPUs (one process each): every rank re-uploads the scene from its host arrays (rtw_scene_update), renders its shard, the
    # int64 buffers are combined with ONE NCCL reduce-scatter, and every rank converts and downloads ITS slice of the image over its
    # own PCIe link into one pinned host buffer shared by the ranks (/dev/shm), which rank 0 reads as the whole picture.
    e2e = e2e_cached = e2e_rgb8 = None
    import ctypes as C

    def host_call(flags=0, out_ptr=None, rgb8=False):
        cfg = rtw.make_cfg(width, height, spp, depth, kernel=kernel, seed=0, device=local_rank, rays_per_lane=args.rays_per_lane, flags=flags)
        d = scene.desc()
        stt = rtw.Stats()
        fn = rtw.lib().rtw_render_rgb8 if rgb8 else rtw.lib().rtw_render
        if fn(C.byref(d), C.byref(cfg), C.c_void_p(out_ptr), C.byref(stt)) != 0:
            raise RuntimeError(rtw.lib().rtw_last_error().decode())
        return stt

    def timed(fn, n):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        barrier()
        ms = (time.perf_counter() - t0) * 1e3 / n
        te = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return te.item()

    if not args.no_e2e:
        scene_bytes = scene.prims.nbytes + scene.mats.nbytes + 168
        e_steps = max(1, min(args.steps, 2))
        if world == 1:
            pinned = torch.empty((height, width, 4), dtype=torch.float32).pin_memory()
            pinned8 = torch.empty((height, width, 3), dtype=torch.uint8).pin_memory()
            parts = {}
            stt_last = {}
            def step_nocache():
                stt = host_call(rtw.FLAG_NO_SCENE_CACHE, pinned.data_ptr())
                stt_last["st"] = stt
                parts.update(flatten_build_upload_ms=stt.h2d_ms, kernel_ms=stt.kernel_ms, d2h_ms=stt.d2h_ms, call_ms=stt.total_ms)
            ms = timed(step_nocache, e_steps)
            e2e = {"value": paths_total / (ms * 1e-3) / 1e6, "unit": "Mpaths/s", "h2d_bytes_per_step": int(scene_bytes), "d2h_bytes_per_step": int(npix * 16),
                   "ms_per_step": ms, "steps": e_steps, "api": "rtw_render (host buffers, RTW_FLAG_NO_SCENE_CACHE: flatten + BVH build + upload every step)", **parts}
            if stt_last["st"].bvh_build_gpu_ms == 0:   # same tree as the timed device-resident render: the very same integers
                assert torch.equal(out_f32.cpu(), pinned), "host-buffer render and device-resident render disagree"
            else:                                      # another (device-built) tree: same closest hits except at exact fp32 ties
                dd = (out_f32.cpu() - pinned).abs()[..., :3] / out_f32.cpu()[..., :3].clamp(min=1e-3)
                assert (dd.amax(dim=2) > 1e-5).float().mean().item() < 2e-3, "host-buffer render (device-built BVH) and device-resident render disagree"
            parts_c = {}
            def step_cached():
                stt = host_call(0, pinned.data_ptr())
                parts_c.update(scene_ms=stt.h2d_ms, kernel_ms=stt.kernel_ms, d2h_ms=stt.d2h_ms, cache_hit=stt.scene_cache_hit)
            ms_c = timed(step_cached, e_steps)
            e2e_cached = {"value": paths_total / (ms_c * 1e-3) / 1e6, "unit": "Mpaths/s", "ms_per_step": ms_c, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(npix * 16),
                          "api": "rtw_render, scene already on the device (hash of the host arrays matches): progressive slices, repeated frames", **parts_c}
            if len(scene.prims) >= 200000:
                # big meshes: where the BVH is built matters end to end.  `e2e` above lets the library choose (device-built linear BVH with a
                # SAH top for renders under 2e9 paths); here the same call with the builder forced either way
                parts["bvh_build_device_ms"] = stt_last["st"].bvh_build_gpu_ms
                for label, fl in (("host_sah", rtw.FLAG_BVH_BUILD_HOST), ("device_lbvh", rtw.FLAG_BVH_BUILD_GPU)):
                    pp = {}
                    def step_forced(fl=fl, pp=pp):
                        stt = host_call(rtw.FLAG_NO_SCENE_CACHE | fl, pinned.data_ptr())
                        pp.update(flatten_build_upload_ms=stt.h2d_ms, bvh_build_device_ms=stt.bvh_build_gpu_ms, kernel_ms=stt.kernel_ms, d2h_ms=stt.d2h_ms)
                    msf = timed(step_forced, e_steps)
                    e2e["builder_" + label] = {"value": paths_total / (msf * 1e-3) / 1e6, "ms_per_step": msf, **pp}
            ms_8 = timed(lambda: host_call(rtw.FLAG_NO_SCENE_CACHE, pinned8.data_ptr(), rgb8=True), e_steps)
            e2e_rgb8 = {"value": paths_total / (ms_8 * 1e-3) / 1e6, "unit": "Mpaths/s", "ms_per_step": ms_8, "h2d_bytes_per_step": int(scene_bytes), "d2h_bytes_per_step": int(npix * 3),
                        "api": "rtw_render_rgb8 (what the drop-in render() calls): write_color on the device, 3 bytes per pixel come back"}
        else:
            # one pinned host image shared by all ranks
            shm_path = f"/dev/shm/rtw_bench_{os.environ.get('MASTER_PORT', '0')}_{npix}.f32"
            if rank == 0:
                with open(shm_path, "wb") as f:
                    f.truncate(npix * 16)
            dist.barrier()
            shared = torch.from_file(shm_path, shared=True, size=npix * 4, dtype=torch.float32)
            registered = torch.cuda.cudart().cudaHostRegister(shared.data_ptr(), npix * 16, 0)
            registered = int(registered) == 0 if not isinstance(registered, tuple) else int(registered[0]) == 0
            ds2 = rtw.DeviceScene(scene, local_rank)
            flat_i64 = accum.view(-1)
            if rows:
                mine_f32 = torch.zeros((local_rows, width, 4), dtype=torch.float32, device=dev)
            else:
                per = (npix * 4 + world - 1) // world
                per += (-per) % 4
                padded = torch.zeros(per * world, dtype=torch.int64, device=dev)
                mine_i64 = torch.zeros(per, dtype=torch.int64, device=dev)
                mine_f32 = torch.zeros(per, dtype=torch.float32, device=dev)
                lo = rank * per
                hi = min(lo + per, npix * 4)

            def step_multi():
                ds2.update(scene)   # host arrays -> HBM (flatten + BVH build + one H2D copy into the existing allocation)
                accum.zero_()
                ds2.render_into(accum, width, height, s_end - s_begin, depth, sample_begin=s_begin, stream_ptr=stream.cuda_stream, seed=0,
                                kernel=kernel, rays_per_lane=args.rays_per_lane, row_tiles=row_tiles)
                if rows:   # no exchange: every rank converts and downloads its own tiles into their rows of the shared image
                    ds2.accum_to_float(accum, mine_f32, local_rows * width, stream_ptr=stream.cuda_stream)
                    img = shared.view(height, width, 4)
                    for k in range(local_rows // args.tile_rows):
                        r0 = (k * world + rank) * args.tile_rows
                        if r0 >= height:
                            break
                        n = min(args.tile_rows, height - r0)
                        img[r0:r0 + n].copy_(mine_f32[k * args.tile_rows:k * args.tile_rows + n], non_blocking=True)
                else:      # ONE collective: reduce-scatter of the int64 sums; every rank owns 1/world of the pixels
                    src = flat_i64
                    if per * world != npix * 4:
                        padded[:npix * 4].copy_(flat_i64)
                        src = padded
                    dist.reduce_scatter_tensor(mine_i64, src, op=dist.ReduceOp.SUM)
                    ds2.accum_to_float(mine_i64, mine_f32, per // 4, stream_ptr=stream.cuda_stream)
                    if hi > lo:
                        shared[lo:hi].copy_(mine_f32[:hi - lo], non_blocking=True)
                torch.cuda.synchronize()
            ms = timed(step_multi, e_steps)
            e2e = {"value": paths_total / (ms * 1e-3) / 1e6, "unit": "Mpaths/s", "h2d_bytes_per_step": int(scene_bytes) * world, "d2h_bytes_per_step": int(npix * 16),
                   "ms_per_step": ms, "steps": e_steps, "shared_pinned_host_image": bool(registered),
                   "api": ("rtw_scene_update + rtw_render_device per rank, " + ("row tiles converted and downloaded by their owners" if rows else
                           "one NCCL reduce-scatter (int64), every rank converts and downloads its slice") + " into one pinned host buffer shared by the ranks")}
            if rank == 0:
                got = shared.view(height, width, 4).clone()
                e2e["host_image_equals_device_image"] = bool(torch.equal(got, out_f32.cpu()))
                assert e2e["host_image_equals_device_image"], "the image assembled in host memory differs from the device-resident result"
            barrier()
            torch.cuda.cudart().cudaHostUnregister(shared.data_ptr())
            del shared
            ds2.close()
            if rank == 0:
                try:
                    os.unlink(shm_path)
                except OSError:
                    pass

    # ---- correctness of what was just timed ----------------------------------------------------------------------------------------
    checks = {}
    if world > 1:
        # rank 0 renders every sample alone: the combined multi-GPU image must be the same integers
        step()      # the timed step once more, so that rank 0 holds the combined int64 buffer again (the e2e leg reused `accum`)
        barrier()
        if rank == 0:
            solo = torch.zeros((height, width, 4), dtype=torch.int64, device=dev)
            ds.render_into(solo, width, height, spp, depth, sample_begin=0, stream_ptr=stream.cuda_stream, seed=0, kernel=kernel, rays_per_lane=args.rays_per_lane)
            solo_f32 = torch.zeros_like(out_f32)
            ds.accum_to_float(solo, solo_f32, npix, stream_ptr=stream.cuda_stream)
            torch.cuda.synchronize()
            checks["image_matches_1gpu"] = bool(torch.equal(solo_f32, out_f32)) and bool(torch.equal(solo, accum_full if rows else accum))
            del solo
        barrier()
        dist.destroy_process_group()
        if rank != 0:
            return 0
        assert checks["image_matches_1gpu"], "the multi-GPU image differs from the one-GPU image"
        # the other ranks are gone: the single-process driver behind the drop-in render() (rtw_render_multi_gpu: one host thread per
        # GPU, peer-memory combine) must produce the same image again, with both splits
        host_img = torch.empty((height, width, 4), dtype=torch.float32).pin_memory()
        ref_img = out_f32.cpu()
        for name, flags in (("inprocess", 0), ("inprocess_rows", rtw.FLAG_SPLIT_ROWS)):
            cfg = rtw.make_cfg(width, height, spp, depth, kernel=kernel, seed=0, device=0, rays_per_lane=args.rays_per_lane, flags=flags,
                               row_tiles=(args.tile_rows, 0, 0) if flags else None)
            d = scene.desc()
            best = None
            for it in range(3):
                stt = rtw.Stats()
                t0 = time.perf_counter()
                if rtw.lib().rtw_render_multi_gpu(C.byref(d), C.byref(cfg), world, C.c_void_p(host_img.data_ptr()), C.byref(stt)) != 0:
                    raise RuntimeError(rtw.lib().rtw_last_error().decode())
                ms = (time.perf_counter() - t0) * 1e3
                if it == 0:
                    checks[name + "_first_call_ms"] = ms   # contexts on the other GPUs + peer access + upload
                else:
                    best = ms if best is None else min(best, ms)
            checks[name + "_matches_1gpu"] = bool(torch.equal(host_img, ref_img))
            checks[name + "_ms"] = best
            checks[name + "_mpaths_per_s"] = paths_total / (best * 1e-3) / 1e6
            checks[name + "_kernel_ms"] = stt.kernel_ms
            assert checks[name + "_matches_1gpu"], f"rtw_render_multi_gpu ({name}) differs from the one-GPU image"
    # config 2 at its full size: the picture that was timed against the reference's own render of the same frame (tests/golden)
    gold = ROOT / "tests" / "golden" / "cover_1080p_1024spp_depth50.npz"
    if args.workload == "cover_1080p_1024spp_depth50" and spp == 1024 and gold.exists():
        g = np.load(gold)
        mine = rtw.quantize(out_f32.cpu().numpy(), spp)
        mse = np.mean((mine.astype(np.float64) - g["rgb"].astype(np.float64)) ** 2)
        mean_ch = out_f32[..., :3].double().mean(dim=(0, 1)).cpu().numpy() / spp
        checks["psnr_vs_reference_render_db"] = float(10 * np.log10(255.0 ** 2 / mse))
        checks["image_mean_z_vs_reference"] = [float(x) for x in (mean_ch - g["mean_ch"]) / (np.sqrt(2.0) * g["se_ch"])]
        checks["reference_render"] = "tests/golden/cover_1080p_1024spp_depth50.npz (oracle/_ref, 1920x1080, 1024 spp, depth 50)"

    # ---- cold end to end of the drop-in: a fresh `rtweekend` process, the reference's own metric (its "Done in" line, render.cpp:188-190)
    e2e_cold = None
    if not args.no_e2e and not args.no_cold:
        cmd = [str(rtw.EXE_PATH), "-w", str(width), "-a", repr(aspect), "-s", str(spp), "-c", str(depth), "-t", "1", "--gpus", str(world)]
        if args.workload.startswith("cover"):
            cmd += ["-n", str(nsqrt)] + ([] if moving else ["--static-spheres"])
        else:
            cmd += ["-l", mesh_path(rtw, wl), "--scene", "mesh-on-ground"]
        if args.kernel != "auto":
            cmd += ["--kernel", args.kernel]
        if rows:
            cmd += ["--split", "rows", "--tile-rows", str(args.tile_rows)]
        runs = []
        for _ in range(2):
            t0 = time.perf_counter()
            with open(os.devnull, "wb") as null:
                pr = subprocess.run(cmd, stdout=null, stderr=subprocess.PIPE)
            wall = (time.perf_counter() - t0) * 1e3
            err = pr.stderr.decode(errors="replace")
            done = re.findall(r"Done in (\d+)ms", err)
            host = re.findall(r"host: (.*)", err)
            kern = re.findall(r"kernel ([0-9.]+) ms", err)
            if pr.returncode != 0 or not done:
                runs.append({"error": err[-300:]})
                continue
            runs.append({"process_wall_ms": wall, "done_in_ms": int(done[-1]), "kernel_ms": float(kern[-1]) if kern else None, "host_breakdown": host[-1] if host else None})
        ok = [r for r in runs if "done_in_ms" in r]
        if ok:
            bestr = min(ok, key=lambda r: r["process_wall_ms"])
            e2e_cold = {"value": paths_total / (bestr["process_wall_ms"] * 1e-3) / 1e6, "unit": "Mpaths/s", **bestr, "runs": len(ok),
                        "what": "fresh process: " + " ".join(cmd[1:]) + " > /dev/null (wall clock of the whole process: exec, scene construction, CUDA contexts, "
                                "flatten + upload, render, combine + download, P3 text); done_in_ms is the reference's own metric, its 'Done in' line"}
        else:
            e2e_cold = {"error": runs[-1].get("error") if runs else "not run"}

    # ---- roofline of the dominant kernel --------------------------------------------------------------------------------
    # Neither HBM nor tensor cores bound this path (scene tables live in shared memory, HBM traffic is the 66 MB
    # accumulation buffer): SURVEY 8(d) names the FP32 FMA pipe.  Algorithmic flops use SURVEY's canonical per-test costs
    # (FMA = 2) times the tests the kernel actually performs, counted by an instrumented pass of the same kernel.
    kinds = scene.prims["kind"]
    radius = np.abs(scene.prims["radius"])
    n_moving = int(((kinds == rtw.RTW_MOVING_SPHERE) & (radius < 100)).sum())
    n_static = int(((kinds == rtw.RTW_SPHERE) & (radius < 100)).sum())
    n_big = int((radius >= 100).sum())
    rays_gpu = rays_total / world  # per launch (per GPU)
    paths_gpu = paths_total / world
    peak_tflops, _ = rtw.fp32_peak(local_rank, 1.0)
    peak_note = "FP32 FMA micro-benchmark (rtw_fp32_peak: scalar FFMA and packed FFMA2 chains, the higher rate) measured in this run; MEASURED_PEAKS.json carries no FP32 figure; nominal %.1f" % NOMINAL_FP32_TFLOPS
    fam = "cover" if args.workload.startswith("cover") else ("dragon" if args.workload.startswith("dragon") else "suzanne")

    def ncu(kernel_key, prefer=""):
        return load_profile(fam, kernel_key, prefer) or (load_profile(fam, kernel_key) if prefer else {})

    def traffic(kernel_key):
        if args.traffic_bytes is not None:
            return args.traffic_bytes
        if not args.workload.startswith(("cover_1080p", "dragon", "suzanne")):
            return None
        # the capture at the benched sample count when there is one, else the 8-spp capture (the traffic does not scale with spp)
        return ncu(kernel_key, f"_{spp}spp").get("traffic")

    def issue_roofline(kernel_key, rays, ms):
        """Instruction-issue roofline: thread instructions per second against SMs x 4 schedulers x 32 lanes x SM clock."""
        n = ncu(kernel_key)
        if not n or "tinst_per_ray" not in n:
            return None
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        mhz = (clocks or {}).get("sm_mhz") or 1965.0
        peak = sm_count * 4 * 32 * mhz * 1e6 / 1e12
        ach = n["tinst_per_ray"] * rays / (ms * 1e-3) / 1e12
        return {"bound": "instruction_issue", "achieved": ach, "peak": peak, "unit": "T thread-instructions/s", "frac": ach / peak,
                "thread_instructions_per_ray": n["tinst_per_ray"], "lanes_per_instruction": n["lanes_per_inst"],
                "issue_active_pct": n.get("issue_active_pct"), "source": n["file"] + " (ncu --set full, one launch; parsed by bench.py)",
                "peak_source": f"{sm_count} SMs x 4 schedulers x 32 lanes x {mhz:.0f} MHz (SM clock sampled during the timed region)"}

    moving_frac = n_moving / max(n_moving + n_static, 1)
    test_flop = FLOP_MOVING_TEST * moving_frac + FLOP_STATIC_TEST * (1 - moving_frac)

    def sweep_roofline(ms, rays, paths, label):
        flop = rays * ((n_static + n_big) * FLOP_STATIC_TEST + n_moving * FLOP_MOVING_TEST + FLOP_SHADE) + (rays - paths) * FLOP_HIT
        ach = flop / (ms * 1e-3) / 1e12
        return {"bound": "fp32_fma", "kernel": label, "achieved": ach, "peak": peak_tflops, "unit": "TFLOP/s", "frac": ach / peak_tflops,
                "traffic": traffic("sweep"), "peak_source": peak_note, "algorithmic_flop_per_launch": flop, "kernel_ms": ms,
                "flop_model": f"rays x (({n_static}+{n_big}) static x 17 + {n_moving} moving x 23 + 80) + hits x 40 (SURVEY 8(d))"}

    if kernel_used == rtw.KERNEL_SPHERES_SMEM:
        roofline = sweep_roofline(kernel_ms, rays_gpu, paths_gpu, "k_render_sweep<2> (K1 shared-memory sphere sweep)")
    else:
        # per-ray work of the BVH kernel from an instrumented low-spp pass (the averages do not depend on spp)
        cs = min(8, s_end - s_begin)
        accum.zero_()
        sst = ds.render_into(accum, width, height, cs, depth, sample_begin=s_begin, stream_ptr=stream.cuda_stream, seed=0, kernel=kernel,
                             rays_per_lane=args.rays_per_lane, want_stats=True, stats=True, row_tiles=row_tiles)
        nodes_pr, tests_pr, tris_pr = sst["node_visits"] / sst["rays"], sst["sphere_tests"] / sst["rays"], sst["tri_tests"] / sst["rays"]
        flop = rays_gpu * (nodes_pr * 24.0 + tests_pr * test_flop + tris_pr * 36.0 + n_big * FLOP_STATIC_TEST + FLOP_SHADE) + (rays_gpu - paths_gpu) * FLOP_HIT
        ach = flop / (kernel_ms * 1e-3) / 1e12
        roofline = {"bound": "fp32_fma", "kernel": kernel_label, "achieved": ach,
                    "peak": peak_tflops, "unit": "TFLOP/s", "frac": ach / peak_tflops, "traffic": traffic("wf" if wavefront else "perlane"), "peak_source": peak_note,
                    "algorithmic_flop_per_launch": flop, "kernel_ms": kernel_ms,
                    "per_ray": {"node_visits": nodes_pr, "sphere_tests": tests_pr, "triangle_tests": tris_pr},
                    "flop_model": "rays x (nodes x 2 boxes x 12 + sphere tests x 17|23 + triangle tests x 36 + big spheres x 17 + 80) + hits x 40 (SURVEY 8(d))",
                    "note": "culling removes ~97% of the sweep's flops, so the flop fraction is low by construction; what limits this kernel is "
                            "instruction issue under divergence: see roofline_issue (thread instructions per second against the issue peak)"}
    roofline_issue = issue_roofline("sweep" if kernel_used == rtw.KERNEL_SPHERES_SMEM else ("wf" if wavefront else "perlane"), rays_gpu, kernel_ms)
    # the SURVEY's FP32-roofline target is defined on the brute-force sweep: measure that kernel too (reduced spp, same scene)
    roofline_sweep = None
    if world == 1 and kernel_used != rtw.KERNEL_SPHERES_SMEM and not scene_has_triangles:
        k1_spp = min(128, spp)
        evs = []
        for it in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            accum.zero_()
            a.record(stream)
            ds.render_into(accum, width, height, k1_spp, depth, sample_begin=0, stream_ptr=stream.cuda_stream, seed=0, kernel=rtw.KERNEL_SPHERES_SMEM)
            b.record(stream)
            evs.append((a, b))
        torch.cuda.synchronize()
        k1_ms = min(a.elapsed_time(b) for a, b in evs[1:])
        frac_spp = k1_spp / spp
        roofline_sweep = sweep_roofline(k1_ms, rays_gpu * frac_spp, paths_gpu * frac_spp, f"k_render_sweep<2> (K1 shared-memory sphere sweep), {k1_spp} spp")
        roofline_sweep["mpaths_per_s"] = paths_gpu * frac_spp / (k1_ms * 1e-3) / 1e6

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline and not scene_has_triangles:
        threads = min(os.cpu_count() or 1, 64)
        cpu_baseline = cpu_reference_run(max(width // 2, 200), aspect, 1, depth, threads)
        if threads > 4:  # also at the reference's default -t 4 (all threads share one unsynchronised mt19937)
            t4 = cpu_reference_run(max(width // 4, 200), aspect, 1, depth, 4)
            cpu_baseline["reference_default_threads"] = {"value": t4["value"], "unit": "Mpaths/s", "cores": 4, "sample": t4["sample"]}
    elif world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_reference_mesh(scene, aspect, depth)

    line = {
        "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "width": width, "height": height, "spp": spp, "max_child_rays": depth, "primitives": int(len(scene.prims)),
                   "parallelism": (f"row tiles of {args.tile_rows} rows interleaved over {world} GPUs, one int64 NCCL gather" if rows
                                   else f"spp-shard x{world}, one int64 NCCL reduce"), "kernel": "spheres_smem (K1)" if kernel_used == rtw.KERNEL_SPHERES_SMEM else ("bvh wavefront (K2w)" if wavefront else "bvh per-lane (K2)"),
                   "l2": "256 MB buffer written between timed iterations (scene tables live in shared memory; accumulation buffer 66 MB)"},
        "mrays_per_s": rays_total / (ms_per_step * 1e-3) / 1e6, "rays_per_path": rays_total / paths_total,
        "e2e": e2e, "e2e_cached": e2e_cached, "e2e_rgb8": e2e_rgb8, "e2e_cold": e2e_cold,
        # kernels of librtw_b200.so launched inside the timed region, all ranks: counted at the launch sites (rtw_kernel_launches), per
        # step one render kernel per GPU + k_accum_to_float (+ k_untile) on rank 0
        "gpu_launches": int(launches_timed), "checks": checks, **{k: v for k, v in checks.items() if k in ("image_matches_1gpu", "inprocess_ms")},
        "clocks": clocks, "roofline": roofline, "roofline_issue": roofline_issue, "roofline_sphere_sweep": roofline_sweep,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
