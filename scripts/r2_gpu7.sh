#!/bin/bash
# K2w with the nodes at an 80-byte stride in shared memory: parity subset, speed at 128 and 1024 spp, bank-conflict counters
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "cover or crossover or determinism or row_tile or full_size or depth_rule" 2>&1 | tail -4
for i in 1 2; do python scripts/profile_render.py --kernel bvh --spp 128 2>&1 | tail -1; done
for i in 1 2; do python scripts/profile_render.py --kernel bvh --spp 1024 2>&1 | tail -1; done
for n in 13 16; do python - <<PY
import importlib,sys
sys.path.insert(0,'.')
rtw=importlib.import_module('raytracing-one-weekend_b200')
sc=rtw.cover_scene($n, 1.7777777777777777)
for i in range(2): acc, st = rtw.render(sc, 1920, 1080, 128, 50)
print("nsqrt $n", len(sc.prims), "prims", round(st['paths']/st['kernel_ms']/1e3,1), "Mpaths/s")
PY
done
python scripts/profile_render.py --kernel bvh --spp 8 > gpurun_out/plain_k2w.log 2>&1 &&
ncu --clock-control none -k regex:k_render -s 1 -c 1 -f -o gpurun_out/prof_k2w_pad --metrics gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed python scripts/profile_render.py --kernel bvh --spp 8 > gpurun_out/ncu_k2w_pad.log 2>&1
ncu -i gpurun_out/prof_k2w_pad.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
for h,u,v in zip(rows[0],rows[1],rows[2]):
    if any(k in h for k in ('l1tex','smsp__','gpu__time')) and ('.sum' in h or 'ratio' in h or 'avg.pct' in h): print(h,v,u)
"
