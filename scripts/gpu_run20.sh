#!/bin/bash
# confirmation of the round's final state: full gpu suite, smoke, the default bench line, ncu of the attention kernel (lazy accumulator)
mkdir -p gpurun_out
{
echo "== full gpu suite"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 --timeout 900 -p no:cacheprovider
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "== bench N=1 K=6"; timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_n1_final.json; tail -c 1500 gpurun_out/bench_n1_final.json
echo "== ncu attention"; timeout 240 ncu --set full --import-source on --clock-control none -k regex:encoder_attention_tc --launch-skip 4 -c 1 -f -o gpurun_out/r02_ncu_attention_lazy python scripts/bench_encoder.py large-v3 8 > gpurun_out/ncu_attention_lazy.log 2>&1; tail -2 gpurun_out/ncu_attention_lazy.log
} > gpurun_out/run20.log 2>&1
tail -30 gpurun_out/run20.log
