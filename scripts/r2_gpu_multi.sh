#!/bin/bash
# Round 2, multi-GPU call:   gpurun --gpus N --timeout 2400 -- 'bash scripts/r2_gpu_multi.sh N [full]'
# N-rank bench (sample split and row split) with the driver-visible identity checks, the single-process driver, cold runs of the
# drop-in executable, and (full) the config-2 scaling points and the config-5 sweep at 1/2/4/8.
N=${1:-2}; FULL=${2:-}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo$N.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -x -k "multi_gpu or concurrent" > gpurun_out/pytest_multi_n$N.log 2>&1; tail -3 gpurun_out/pytest_multi_n$N.log
run() {  # gpus, log name, extra args
  local n=$1 name=$2; shift 2
  if [ "$n" = 1 ]; then timeout 1500 python bench.py --gpus 1 "$@" > gpurun_out/$name.log 2>&1
  else timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n "$@" > gpurun_out/$name.log 2>&1; fi
  echo "rc=$?" >> gpurun_out/$name.log
  grep -E '^\{"metric"' gpurun_out/$name.log | python -c "
import json,sys
for ln in sys.stdin:
    d=json.loads(ln)
    print('$name', 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', d['e2e'] and round(d['e2e']['value']), 'e2e_ms', d['e2e'] and round(d['e2e']['ms_per_step'],2), 'cold', d.get('e2e_cold') and {k: d['e2e_cold'].get(k) for k in ('process_wall_ms','done_in_ms','host_breakdown')}, 'checks', d.get('checks'), 'launches', d.get('gpu_launches'))
" || tail -5 gpurun_out/$name.log
}
run $N bench_n$N --steps 3 --warmup 3
run $N bench_n${N}_rows --steps 3 --warmup 3 --split rows --no-cpu-baseline --no-cold
EXE=./raytracing-one-weekend_b200/rtweekend
for g in 1 $N; do for i in 1 2; do
  t0=$(date +%s%N); RTW_TRACE=1 $EXE -w 1920 -a 1.7777777777777777 -s 1024 -c 50 --gpus $g > gpurun_out/cover_${g}gpu.ppm 2> gpurun_out/cover_${g}gpu_$i.err; t1=$(date +%s%N)
  echo "process wall $(( (t1 - t0) / 1000000 )) ms" >> gpurun_out/cover_${g}gpu_$i.err; grep -v "^Started\|^$" gpurun_out/cover_${g}gpu_$i.err | tr '\n' ' '; echo
done; done
cmp gpurun_out/cover_1gpu.ppm gpurun_out/cover_${N}gpu.ppm && echo "1-GPU and $N-GPU PPM identical"
$EXE -w 1920 -a 1.7777777777777777 -s 1024 -c 50 --gpus $N --split rows > gpurun_out/cover_rows.ppm 2> gpurun_out/cover_rows.err; tail -2 gpurun_out/cover_rows.err | tr '\n' ' '; echo
cmp gpurun_out/cover_1gpu.ppm gpurun_out/cover_rows.ppm && echo "1-GPU and $N-GPU row-split PPM identical"
$EXE -w 400 -s 64 --gpus $N --checkpoint /tmp/ck_multi.bin --checkpoint-every 16 > gpurun_out/ck_multi.ppm 2> gpurun_out/ck_multi.err; tail -3 gpurun_out/ck_multi.err | tr '\n' ' '; echo
$EXE -w 400 -s 64 --gpus 1 --checkpoint /tmp/ck_one.bin --checkpoint-every 16 > gpurun_out/ck_one.ppm 2>/dev/null; cmp gpurun_out/ck_one.ppm gpurun_out/ck_multi.ppm && echo "progressive: 1-GPU and $N-GPU PPM identical"
md5sum gpurun_out/*.ppm; rm -f gpurun_out/*.ppm
if [ -n "$FULL" ]; then   # config 5 (4K x 4096 spp) at this GPU count: one point of the strong-scaling sweep (the other counts run on smaller boxes)
  run $N bench_4k_n$N --steps 2 --warmup 3 --workload cover_4k_4096spp_depth50 --no-cpu-baseline --no-cold
fi
