#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/run4.log
{
echo "== depth 2, K=6"; timeout 600 python bench.py --value-only --steps 6 --warmup 1 --depth 2
echo "== depth 3, K=6"; timeout 600 python bench.py --value-only --steps 6 --warmup 1 --depth 3
echo "== depth 2, K=6, 192 rows"; timeout 600 python bench.py --value-only --steps 6 --warmup 1 --depth 2 --slots 192
echo "== depth 2, K=6, 320 rows"; timeout 600 python bench.py --value-only --steps 6 --warmup 1 --depth 2 --slots 320
echo "== depth 2, K=6, serial encoder"; OH_ASYNC_ENC=0 timeout 600 python bench.py --value-only --steps 6 --warmup 1 --depth 2
} > $o 2>&1
tail -30 $o
