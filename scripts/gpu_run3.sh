#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/run3.log
{
echo "== decode step 256 rows, pos 100 (diag lib, elimination)"; timeout 300 python scripts/bench_decode_step.py large-v3 256 100
echo "== decode step 128 rows"; timeout 200 python scripts/bench_decode_step.py large-v3 128 100 quick
echo "== dual 256"; OH_DUAL=256 timeout 200 python scripts/bench_decode_step.py large-v3 256 100 quick
echo "== dual 256 + xattn occ 1"; OH_DUAL=256 OH_XATTN_OCC=1 timeout 200 python scripts/bench_decode_step.py large-v3 256 100 quick
echo "== xattn occ 1 alone"; OH_XATTN_OCC=1 timeout 200 python scripts/bench_decode_step.py large-v3 256 100 quick
echo "== encoder 32 windows"; timeout 200 python scripts/bench_encoder.py large-v3 32
echo "== bench value-only, starve logic"; timeout 500 python bench.py --value-only --steps 2 --warmup 1
echo "== bench value-only, starve 1"; OH_STARVE=1 timeout 500 python bench.py --value-only --steps 2 --warmup 1
echo "== bench value-only, 128 rows"; timeout 500 python bench.py --value-only --steps 2 --warmup 1 --slots 128
} > $o 2>&1
tail -60 $o
