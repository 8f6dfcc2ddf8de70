#!/bin/bash
# final kernels on N GPUs: config 2 (optional) and config 5 (4K x 4096 spp) through the N-rank bench
N=${1:-4}; WHAT=${2:-both}
mkdir -p gpurun_out
run() {
  local name=$1; shift
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" > gpurun_out/$name.log 2>&1
  echo "rc=$?" >> gpurun_out/$name.log
  grep -E '^\{"metric"' gpurun_out/$name.log | python -c "
import json,sys
for ln in sys.stdin:
    d=json.loads(ln)
    c=d.get('checks') or {}
    print('$name', 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', d['e2e'] and round(d['e2e']['value']), 'identical', c.get('image_matches_1gpu'), c.get('inprocess_matches_1gpu'), 'inprocess', c.get('inprocess_mpaths_per_s'))
" || tail -5 gpurun_out/$name.log
}
[ "$WHAT" != 4k ] && run bench_n$N --steps 3 --warmup 3 --no-cold
run bench_4k_n$N --steps 2 --warmup 3 --workload cover_4k_4096spp_depth50 --no-cpu-baseline --no-cold
