#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/run12.log
{
echo "== baseline 256"; timeout 200 python scripts/bench_decode_step.py large-v3 256 100 quick
echo "== dual 256 + xattn stages 3"; OH_DUAL=256 OH_XATTN_STAGES=3 timeout 200 python scripts/bench_decode_step.py large-v3 256 100 quick
echo "== dual 256 + xattn stages 2"; OH_DUAL=256 OH_XATTN_STAGES=2 timeout 200 python scripts/bench_decode_step.py large-v3 256 100 quick
echo "== xattn stages 2 alone"; OH_XATTN_STAGES=2 timeout 200 python scripts/bench_decode_step.py large-v3 256 100 quick
} > $o 2>&1
tail -40 $o
