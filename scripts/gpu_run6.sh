#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/run6.log
{
echo "== quantised model test"; timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -k "quantised or worker_pool" --timeout 600 -p no:cacheprovider
echo "== xattn stages 3"; OH_XATTN_STAGES=3 timeout 200 python scripts/bench_decode_step.py large-v3 256 100 quick
echo "== xattn stages 4"; OH_XATTN_STAGES=4 timeout 200 python scripts/bench_decode_step.py large-v3 256 100 quick
echo "== encoder compact 96 alone"; OH_ENC_COMPACT=96 OH_ENC_COMPACT_ALL=1 timeout 200 python scripts/bench_encoder.py large-v3 32
echo "== encoder compact 128 alone"; OH_ENC_COMPACT=128 OH_ENC_COMPACT_ALL=1 timeout 200 python scripts/bench_encoder.py large-v3 32
echo "== bench K=4 compact96 + xattn3, prio lo"; OH_ENC_COMPACT=96 OH_XATTN_STAGES=3 OH_ENC_PRIO=lo timeout 600 python bench.py --value-only --steps 4 --warmup 1
echo "== bench K=4 compact96 + xattn3, prio hi"; OH_ENC_COMPACT=96 OH_XATTN_STAGES=3 timeout 600 python bench.py --value-only --steps 4 --warmup 1
echo "== bench K=4 compact128 + xattn3, prio lo"; OH_ENC_COMPACT=128 OH_XATTN_STAGES=3 OH_ENC_PRIO=lo timeout 600 python bench.py --value-only --steps 4 --warmup 1
echo "== bench K=4 xattn3 only"; OH_XATTN_STAGES=3 timeout 600 python bench.py --value-only --steps 4 --warmup 1
} > $o 2>&1
tail -50 $o
