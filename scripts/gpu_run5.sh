#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/run5.log
{
echo "== gpu tests: stream + beam + engine"; timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_engine.py tests/test_gpu_parity.py -m gpu -q --maxfail=5 --timeout 600 -p no:cacheprovider
echo "== trace run depth 2, K=4"; OH_SCHED_TRACE=gpurun_out/sched_trace.csv timeout 600 python bench.py --value-only --steps 4 --warmup 1 --depth 2
echo "== beam bench value-only K=1"; timeout 900 python bench.py --value-only --steps 1 --warmup 1 --strategy beam --batch 128
} > $o 2>&1
tail -40 $o
