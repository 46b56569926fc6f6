#!/bin/bash
# lazy-rescale attention accumulator (OH_ATTN_O=lazy): self-test, encoder timing, parity subset
mkdir -p gpurun_out
{
echo "== encoder, default"; timeout 200 python scripts/bench_encoder.py large-v3 32
echo "== encoder, lazy accumulator"; OH_ATTN_O=lazy timeout 200 python scripts/bench_encoder.py large-v3 32
echo "== parity subset, lazy"; OH_ATTN_O=lazy timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -q --maxfail=5 --timeout 600 -p no:cacheprovider
} > gpurun_out/run19.log 2>&1
tail -30 gpurun_out/run19.log
