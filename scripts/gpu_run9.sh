#!/bin/bash
# 2-GPU validation: in-process worker-per-GPU tests (per-device kernel set-up), 2-rank bench
mkdir -p gpurun_out
o=gpurun_out/run9.log
{
nvidia-smi -L
echo "== worker pool / host test / overlap streaming on 2 GPUs"; timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_resample.py -m gpu -q --timeout 600 -p no:cacheprovider
echo "== frontend"; timeout 200 python scripts/bench_frontend.py large-v3 256
echo "== bench 2 GPUs value-only K=3"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --value-only --steps 3 --warmup 1
} > $o 2>&1
tail -30 $o
