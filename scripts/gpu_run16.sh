#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/run16.log
{
echo "== encoder, P through tensor memory"; OH_ATTN_P=tmem timeout 300 python scripts/bench_encoder.py large-v3 32
echo "== encoder, P through shared memory"; timeout 300 python scripts/bench_encoder.py large-v3 32
echo "== parity subset with P through tensor memory"; OH_ATTN_P=tmem timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider -k "16bit or tcgen05 or f16"
} > $o 2>&1
tail -30 $o
