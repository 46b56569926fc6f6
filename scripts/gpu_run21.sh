#!/bin/bash
# split-K epilogue restructure (10 warps, batched DSMEM loads, residual prefetch) + self-attention occupancy variants
mkdir -p gpurun_out
{
echo "== decode step 256 rows, pos 100, default"; timeout 200 python scripts/bench_decode_step.py large-v3 256 100 quick
echo "== OH_SELF_OCC=5"; OH_SELF_OCC=5 timeout 200 python scripts/bench_decode_step.py large-v3 256 100 quick
echo "== OH_SELF_OCC=6"; OH_SELF_OCC=6 timeout 200 python scripts/bench_decode_step.py large-v3 256 100 quick
echo "== decode step 128 rows"; timeout 200 python scripts/bench_decode_step.py large-v3 128 100 quick
echo "== parity subset"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q --maxfail=5 --timeout 600 -p no:cacheprovider -k "16bit or tcgen05 or f16 or config3 or config2"
} > gpurun_out/run21.log 2>&1
tail -40 gpurun_out/run21.log
