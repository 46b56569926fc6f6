#!/bin/bash
# C++ worker pool with a drain kept in flight on the engine's stream
mkdir -p gpurun_out
timeout 80 python -m pytest "tests/test_gpu_engine.py::test_cpp_host_mirror_gpu" -m gpu -q --timeout 70 -p no:cacheprovider > gpurun_out/run30.log 2>&1
tail -15 gpurun_out/run30.log
