#!/bin/bash
# Round 2, final kernels (32-warp K2w tier, FFMA2 sweep, SAH subtrees in the device build): full parity suite + smoke + headline bench + mesh benches + config 5 at N=1 + ncu evidence (launch list of the bench command, full captures)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=6 ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -16 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log; tail -2 gpurun_out/bench.log | cut -c1-800
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-400
for w in cover_default_200x133_20spp_depth20 suzanne_on_ground_1080p_256spp; do
  timeout 900 python bench.py --steps 3 --warmup 3 --workload $w --no-cold > gpurun_out/bench_$w.log 2>&1; tail -1 gpurun_out/bench_$w.log | cut -c1-500
done
timeout 1200 python bench.py --steps 2 --warmup 3 --workload cover_4k_4096spp_depth50 --no-cpu-baseline --no-cold > gpurun_out/bench_4k_n1.log 2>&1; tail -1 gpurun_out/bench_4k_n1.log | cut -c1-500
# ncu: launch list of the bench command (reduced spp: ncu serialises and replays), then full captures
B="python bench.py --steps 2 --warmup 3 --spp 64 --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv $B > gpurun_out/ncu_launch.log 2>&1
prof() {  # name, then the arguments of scripts/profile_render.py
  local name=$1; shift
  python scripts/profile_render.py "$@" > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_render -s 1 -c 1 -f -o gpurun_out/prof_$name python scripts/profile_render.py "$@" > gpurun_out/ncu_$name.log 2>&1
}
prof k2w --kernel bvh --spp 8
prof k2_perlane --kernel bvh-perlane --spp 8
prof k1 --kernel spheres --spp 8
prof k2_suzanne --kernel bvh --scene suzanne --spp 8 --depth 20
python -c "
import importlib,ctypes as C,sys
sys.path.insert(0,'.')
rtw=importlib.import_module('raytracing-one-weekend_b200')
n=C.c_longlong(0); rtw.host().rtwh_make_mesh(b'assets/suzanne.obj', b'/tmp/standin5.obj', 5, 20221018, 0.08, C.byref(n)); print('tris', n.value)
"
prof k2_dragon --kernel bvh --scene /tmp/standin5.obj --spp 4 --depth 20
timeout 900 python bench.py --steps 3 --warmup 3 --workload dragon_standin_1080p_256spp --no-cold > gpurun_out/bench_dragon_standin.log 2>&1; tail -1 gpurun_out/bench_dragon_standin.log | cut -c1-500
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_bench.csv
python scripts/profile_render.py --kernel bvh --spp 1024 > gpurun_out/plain_k2w_1024spp.log 2>&1 &&
ncu --clock-control none -k regex:k_render -s 1 -c 1 -f -o gpurun_out/prof_k2w_1024spp --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,lts__t_sectors.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum,smsp__inst_executed_op_global_red.sum python scripts/profile_render.py --kernel bvh --spp 1024 > gpurun_out/ncu_k2w_1024spp.log 2>&1
tail -3 gpurun_out/ncu_k2w_1024spp.log
