#!/bin/bash
# remaining fast GPU suites on the final library (engine / pool / host mirror, resampler, VAD, level processing)
mkdir -p gpurun_out
{
timeout 170 python -m pytest tests/test_gpu_engine.py tests/test_resample.py tests/test_vad.py tests/test_preprocess.py -m gpu -q --maxfail=5 --timeout 150 -p no:cacheprovider
} > gpurun_out/run27.log 2>&1
tail -12 gpurun_out/run27.log
