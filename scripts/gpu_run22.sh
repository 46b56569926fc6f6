#!/bin/bash
# confirmation after the split-K epilogue change: full gpu suite, smoke, default bench line
mkdir -p gpurun_out
{
echo "== full gpu suite"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 --timeout 900 -p no:cacheprovider
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "== bench N=1 K=6"; timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_n1_final2.json; tail -c 600 gpurun_out/bench_n1_final2.json
} > gpurun_out/run22.log 2>&1
tail -30 gpurun_out/run22.log
