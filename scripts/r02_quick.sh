#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k "deepsea or other_conv or fused or long_first or non_strand" > gpurun_out/t_quick.log 2>&1; echo "quick tests exit $?"
tail -15 gpurun_out/t_quick.log
