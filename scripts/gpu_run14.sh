#!/bin/bash
# 8-GPU beam-search line (BASELINE config 4's decode policy): value-only
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --strategy beam --value-only --steps 2 --warmup 1 --batch 256 > gpurun_out/r02_bench_beam_large-v3_8gpu.json 2> gpurun_out/bench14.err
echo "rc=$?"; tail -c 1500 gpurun_out/r02_bench_beam_large-v3_8gpu.json; tail -3 gpurun_out/bench14.err
