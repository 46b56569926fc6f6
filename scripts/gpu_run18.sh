#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/run18.log
{
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout 900 -p no:cacheprovider
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "== encoder"; timeout 300 python scripts/bench_encoder.py large-v3 32
echo "== bench N=1 value-only K=10"; timeout 1200 python bench.py --value-only --steps 10 --warmup 2
} > $o 2>&1
tail -30 $o
