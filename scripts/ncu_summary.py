#!/usr/bin/env python
"""Summarise an .ncu-rep (one launch) into the handful of metrics the design argues from.
    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [gpurun_out/plain_run.log] > profiles/r02_xxx.txt
The optional second argument is the log of the same command run WITHOUT ncu (scripts/profile_render.py): its `# launch:` line (paths and
rays of the profiled launch) is copied into the summary so that bench.py can turn instruction counts into per-ray figures."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
KEYS = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed", "sm__sass_thread_inst_executed_op_ffma_pred_on.sum.peak_sustained",
    "sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum.per_second", "lts__t_bytes.sum.per_second",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors.sum", "lts__t_requests_srcunit_tex_op_red.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
]
print(f"# {rep}")
if len(sys.argv) > 2:
    for ln in open(sys.argv[2]):
        if ln.startswith("# launch:"):
            print(ln.rstrip())
for k in KEYS:
    if k in m:
        print(f"{k:85s} {m[k][0]:>22s} {m[k][1]}")
print("# warp stall reasons, cycles per issued instruction")
st = sorted(((float(v[0]), k) for k, v in m.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")), reverse=True)
for v, k in st[:10]:
    print(f"{k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:30s} {v:8.3f}")
