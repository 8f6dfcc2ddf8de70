#!/bin/bash
mkdir -p gpurun_out
python scripts/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1; cat gpurun_out/e2e_probe.log
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_gpu" > gpurun_out/pytest_mgpu.log 2>&1; tail -3 gpurun_out/pytest_mgpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n2.log
tail -3 gpurun_out/bench_n2.log
timeout 600 ./raytracing-one-weekend_b200/rtweekend -w 1920 -a 1.7777777777777777 -s 1024 -c 50 --gpus 2 > gpurun_out/cover_2gpu.ppm 2> gpurun_out/cover_2gpu.err; tail -3 gpurun_out/cover_2gpu.err
timeout 600 ./raytracing-one-weekend_b200/rtweekend -w 1920 -a 1.7777777777777777 -s 1024 -c 50 --gpus 1 > gpurun_out/cover_1gpu.ppm 2> gpurun_out/cover_1gpu.err; tail -3 gpurun_out/cover_1gpu.err
cmp gpurun_out/cover_1gpu.ppm gpurun_out/cover_2gpu.ppm && echo "1-GPU and 2-GPU PPM identical"
md5sum gpurun_out/*.ppm; rm -f gpurun_out/cover_2gpu.ppm; gzip -f gpurun_out/cover_1gpu.ppm
