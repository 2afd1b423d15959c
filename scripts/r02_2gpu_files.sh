#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_analyze.py -x -q -m gpu -p no:cacheprovider -k two_processes > gpurun_out/t_2gpu.log 2>&1; echo "2-gpu vep test exit $?"
tail -3 gpurun_out/t_2gpu.log
