#!/bin/bash
# final kernels on N GPUs: in-process multi-GPU tests + the N-rank bench with its identity checks (sample split)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "multi_gpu or concurrent" > gpurun_out/pytest_multi_n$N.log 2>&1; tail -3 gpurun_out/pytest_multi_n$N.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.log 2>&1
echo "rc=$?" >> gpurun_out/bench_n$N.log
grep -E '^\{"metric"' gpurun_out/bench_n$N.log | python -c "
import json,sys
for ln in sys.stdin:
    d=json.loads(ln)
    print('n$N', 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', d['e2e'] and round(d['e2e']['value']), 'cold', d.get('e2e_cold') and d['e2e_cold'].get('done_in_ms'), 'checks', d.get('checks'), 'launches', d.get('gpu_launches'))
" || tail -5 gpurun_out/bench_n$N.log
