#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:encoder_attention_tc --launch-skip 4 -c 1 -f -o gpurun_out/r02_ncu_attention python scripts/bench_encoder.py large-v3 8 > gpurun_out/ncu_attention.log 2>&1
tail -3 gpurun_out/ncu_attention.log
OH_ATTN_P=tmem timeout 600 ncu --set full --import-source on --clock-control none -k regex:encoder_attention_tc --launch-skip 4 -c 1 -f -o gpurun_out/r02_ncu_attention_ptmem python scripts/bench_encoder.py large-v3 8 > gpurun_out/ncu_attention2.log 2>&1
tail -3 gpurun_out/ncu_attention2.log
