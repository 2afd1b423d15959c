#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -x -q -m gpu -p no:cacheprovider > gpurun_out/t_quick.log 2>&1; echo "exit $?"
tail -5 gpurun_out/t_quick.log
