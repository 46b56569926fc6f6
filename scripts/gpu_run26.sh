#!/bin/bash
# last check of the round on the final library: smoke + the micro-model parity suite
mkdir -p gpurun_out
{
echo "== smoke"; timeout 150 python -c "import __graft_entry__ as g; g.smoke()"
echo "== parity (micro model) + stream"; timeout 120 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stream.py -m gpu -q --maxfail=5 --timeout 100 -p no:cacheprovider
} > gpurun_out/run26.log 2>&1
tail -12 gpurun_out/run26.log
