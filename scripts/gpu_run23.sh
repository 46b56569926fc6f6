#!/bin/bash
# refresh of the decode-kernel ncu capture at 256 rows (the split-K kernel changed its block size and epilogue)
mkdir -p gpurun_out
{
echo "== decode step plain"; timeout 200 python scripts/ncu_decode_step.py large-v3 256 100 4
echo "== decode step ncu full"; timeout 420 ncu --clock-control none --set full --import-source on -k regex:"cross_attention_decode_bulk|gemm_splitk|self_attention_decode_warp|layernorm_kernel|logits_step|embed_kernel" --launch-skip 800 -c 24 -f -o gpurun_out/r02_ncu_decode_v2 python scripts/ncu_decode_step.py large-v3 256 100 6 > gpurun_out/ncu_decode_v2.log 2>&1; tail -3 gpurun_out/ncu_decode_v2.log
} > gpurun_out/run23.log 2>&1
tail -12 gpurun_out/run23.log
