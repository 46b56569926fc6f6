#!/bin/bash
# ncu evidence for profiles/ (each target ran without ncu first)
mkdir -p gpurun_out
o=gpurun_out/run8.log
NCU="ncu --clock-control none"
{
echo "== frontend plain"; timeout 300 python scripts/bench_frontend.py large-v3 256
echo "== frontend ncu full"; timeout 600 $NCU --set full --import-source on -k regex:"mel_frames|resample_sinc" -c 4 -f -o gpurun_out/r02_ncu_frontend python scripts/bench_frontend.py large-v3 256 > gpurun_out/ncu_frontend.log 2>&1; tail -3 gpurun_out/ncu_frontend.log
echo "== decode step plain"; timeout 300 python scripts/ncu_decode_step.py large-v3 256 100 4
echo "== decode step ncu full"; timeout 900 $NCU --set full --import-source on -k regex:"cross_attention_decode_bulk|gemm_splitk|self_attention_decode_warp|layernorm_kernel|logits_step|embed_kernel" --launch-skip 800 -c 24 -f -o gpurun_out/r02_ncu_decode python scripts/ncu_decode_step.py large-v3 256 100 6 > gpurun_out/ncu_decode.log 2>&1; tail -3 gpurun_out/ncu_decode.log
echo "== bench launch list"; timeout 900 $NCU --metrics gpu__time_duration.sum --launch-skip 30000 -c 8000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --value-only --steps 1 --warmup 1 --batch 320 > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log; wc -l gpurun_out/r02_launches_bench.csv
} > $o 2>&1
tail -30 $o
ls -la gpurun_out/*.ncu-rep
