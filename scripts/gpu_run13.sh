#!/bin/bash
# 2-GPU run of the full bench line (N > 1 code paths: all-reduced counters, record gather, strong-scaling leg with the N = 1 reference)
mkdir -p gpurun_out
timeout 1400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_bench_large-v3_2gpu.json 2> gpurun_out/bench13.err
echo "rc=$?"; tail -c 2500 gpurun_out/r02_bench_large-v3_2gpu.json; tail -5 gpurun_out/bench13.err
