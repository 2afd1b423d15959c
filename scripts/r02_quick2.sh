#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_score_parity.py tests/test_gpu_golden.py -x -q -m gpu -p no:cacheprovider -k "deepsea or fused or ahead or score or register or golden" > gpurun_out/t_quick.log 2>&1; echo "quick tests exit $?"
tail -3 gpurun_out/t_quick.log
bash scripts/r02_tcb.sh "SB_X=1" "SB_CONV1_AHEAD=0"
