#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/run11.log
{
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 900 -p no:cacheprovider
echo "== bench N=1 full"; timeout 1500 python bench.py --steps 6 --warmup 2 > gpurun_out/r02_bench_large-v3_1gpu.json 2> gpurun_out/bench11.err; echo "rc=$?"; tail -c 1500 gpurun_out/r02_bench_large-v3_1gpu.json; tail -5 gpurun_out/bench11.err
echo "== bench beam N=1 value-only"; timeout 900 python bench.py --value-only --steps 2 --warmup 1 --strategy beam --batch 256
echo "== reference arm"; OH_REF_BUDGET_S=60 timeout 400 python bench.py --impl reference --steps 2 --warmup 1
} > $o 2>&1
tail -40 $o
