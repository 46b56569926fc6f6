#!/bin/bash
# level processing against the committed golden fixture
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_preprocess.py -m gpu -q --maxfail=5 --timeout 100 -p no:cacheprovider > gpurun_out/run28.log 2>&1
tail -8 gpurun_out/run28.log
