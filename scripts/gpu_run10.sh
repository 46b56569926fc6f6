#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/run10.log
{
echo "== debug overlap"; timeout 300 python scripts/debug_overlap.py
echo "== engine tests"; timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_stream.py -m gpu -q --timeout 600 -p no:cacheprovider
} > $o 2>&1
tail -60 $o
