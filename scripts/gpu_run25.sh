#!/bin/bash
# level-processing kernels (preprocess.cu): parity against the oracle and the reference's unit tests, host mirrors, timing
mkdir -p gpurun_out
{
echo "== preprocess tests"; timeout 240 python -m pytest tests/test_preprocess.py "tests/test_gpu_engine.py::test_cpp_host_mirror_gpu" -m gpu -q --maxfail=5 --timeout 200 -p no:cacheprovider
echo "== timing"; timeout 100 python scripts/bench_preprocess.py 64 30
} > gpurun_out/run25.log 2>&1
tail -40 gpurun_out/run25.log
