#!/bin/bash
# bench.py set-up path after threading the synthetic-audio generation (small model, tiny run)
mkdir -p gpurun_out
timeout 60 python bench.py --value-only --steps 1 --warmup 1 --batch 64 --arch small > gpurun_out/run31.log 2>&1
tail -5 gpurun_out/run31.log
