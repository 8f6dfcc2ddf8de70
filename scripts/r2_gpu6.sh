#!/bin/bash
# Round 2: full parity suite + smoke + headline bench + mesh benches + config 5 at N=1 + ncu evidence (launch list of the bench command, full captures)
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
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_bench.csv
