#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/run15.log
{
echo "== encoder 32 windows (2 softmax threads per row)"; timeout 300 python scripts/bench_encoder.py large-v3 32
echo "== parity: 16-bit modes, configs"; timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q --timeout 900 -p no:cacheprovider
} > $o 2>&1
tail -30 $o
