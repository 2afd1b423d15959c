#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k "tc_gemm_selftest" > gpurun_out/t_pair.log 2>&1; echo "gemm selftest exit $?"; tail -4 gpurun_out/t_pair.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k "deepsea or other_conv or fused" > gpurun_out/t_pair2.log 2>&1; echo "deepsea tests exit $?"; tail -4 gpurun_out/t_pair2.log
bash scripts/r02_tcb.sh "SB_X=1" "SB_TC_PAIR=0" "SB_X=2" "SB_TC_PAIR=0"
