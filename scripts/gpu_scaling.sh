#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
for n in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/bench_n$n.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n$n.log
  tail -2 gpurun_out/bench_n$n.log | cut -c1-700
done
timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | cut -c1-300
# the row-tile alternative of SURVEY 8(e) at the same size
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 8 --steps 3 --warmup 3 --split rows --no-cpu-baseline > gpurun_out/bench_n8_rows.log 2>&1; tail -1 gpurun_out/bench_n8_rows.log | cut -c1-700
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 2 --warmup 3 --workload cover_4k_4096spp_depth50 > gpurun_out/bench_4k_n8.log 2>&1; tail -1 gpurun_out/bench_4k_n8.log | cut -c1-700
timeout 600 ./raytracing-one-weekend_b200/rtweekend -w 1920 -a 1.7777777777777777 -s 1024 -c 50 --gpus 8 > gpurun_out/cover_8gpu.ppm 2> gpurun_out/cover_8gpu.err; tail -3 gpurun_out/cover_8gpu.err
timeout 600 ./raytracing-one-weekend_b200/rtweekend -w 1920 -a 1.7777777777777777 -s 1024 -c 50 --gpus 1 > gpurun_out/cover_1gpu.ppm 2> gpurun_out/cover_1gpu.err
cmp gpurun_out/cover_1gpu.ppm gpurun_out/cover_8gpu.ppm && echo "1-GPU and 8-GPU PPM identical"
timeout 600 ./raytracing-one-weekend_b200/rtweekend -w 1920 -a 1.7777777777777777 -s 1024 -c 50 --gpus 8 --split rows > gpurun_out/cover_8gpu_rows.ppm 2> gpurun_out/cover_8gpu_rows.err; tail -2 gpurun_out/cover_8gpu_rows.err
cmp gpurun_out/cover_1gpu.ppm gpurun_out/cover_8gpu_rows.ppm && echo "1-GPU and 8-GPU row-split PPM identical"
md5sum gpurun_out/*.ppm; rm -f gpurun_out/cover_8gpu.ppm gpurun_out/cover_1gpu.ppm gpurun_out/cover_8gpu_rows.ppm
