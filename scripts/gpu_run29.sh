#!/bin/bash
# worker pool with drains kept in flight on the engine's stream (oh_stream_submit / oh_stream_wait)
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_engine.py -m gpu -q --maxfail=5 --timeout 90 -p no:cacheprovider > gpurun_out/run29.log 2>&1
tail -15 gpurun_out/run29.log
