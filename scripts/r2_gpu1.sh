#!/bin/bash
# Round 2, first GPU call (one GPU): parity suite, smoke, bench (ours + reference arm), cold-start probes, work-unit sweep, ncu captures of
# the headline kernel.   gpurun --timeout 2400 -- 'bash scripts/r2_gpu1.sh'
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt; nproc >> gpurun_out/gpus.txt
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -25 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log; tail -2 gpurun_out/bench.log | cut -c1-3000
# cold start of the drop-in: does restricting the visible devices help?
EXE=./raytracing-one-weekend_b200/rtweekend
for i in 1 2; do $EXE -w 1920 -a 1.7777777777777777 -s 1024 -c 50 > /dev/null 2> gpurun_out/cold_all_$i.err; tail -3 gpurun_out/cold_all_$i.err | tr '\n' ' '; echo; done
for i in 1 2; do CUDA_VISIBLE_DEVICES=0 $EXE -w 1920 -a 1.7777777777777777 -s 1024 -c 50 > /dev/null 2> gpurun_out/cold_vis0_$i.err; tail -3 gpurun_out/cold_vis0_$i.err | tr '\n' ' '; echo; done
( time $EXE -w 1920 -a 1.7777777777777777 -s 1024 -c 50 --format p6 > /dev/null ) 2>&1 | tail -6 | tr '\n' ' '; echo
# samples per work unit (tail vs fetch overhead), 128 spp (the 8-GPU shard) and 1024 spp
for su in 1 2 4 8; do RTW_SU=$su python scripts/profile_render.py --kernel bvh --spp 128 2>&1 | tail -1 | sed "s/^/su=$su /"; done
for su in 4 8 16; do RTW_SU=$su python scripts/profile_render.py --kernel bvh --spp 1024 2>&1 | tail -1 | sed "s/^/su=$su /"; done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-600
# ncu: full capture of the headline kernel at 8 spp, selected counters at the benched 1024 spp
prof() {  # name, then the arguments of scripts/profile_render.py
  local name=$1; shift
  python scripts/profile_render.py "$@" > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_render -s 1 -c 1 -f -o gpurun_out/prof_$name python scripts/profile_render.py "$@" > gpurun_out/ncu_$name.log 2>&1
}
prof k2w --kernel bvh --spp 8
python scripts/profile_render.py --kernel bvh --spp 1024 > gpurun_out/plain_k2w_1024spp.log 2>&1 &&
ncu --clock-control none -k regex:k_render -s 1 -c 1 -f -o gpurun_out/prof_k2w_1024spp --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,lts__t_sectors.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum,smsp__inst_executed_op_global_red.sum python scripts/profile_render.py --kernel bvh --spp 1024 > gpurun_out/ncu_k2w_1024spp.log 2>&1
tail -3 gpurun_out/ncu_k2w_1024spp.log
ls -la gpurun_out/*.ncu-rep
