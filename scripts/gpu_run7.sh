#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/run7.log
{
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 900 -p no:cacheprovider
echo "== bench K=4 value-only"; OH_SCHED_TRACE=gpurun_out/sched_trace7.csv timeout 600 python bench.py --value-only --steps 4 --warmup 1
echo "== bench K=4 value-only 224 rows"; timeout 600 python bench.py --value-only --steps 4 --warmup 1 --slots 224
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
} > $o 2>&1
tail -40 $o
