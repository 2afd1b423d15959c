#!/bin/bash
# Runs the GPU parity tests in separate processes (a trapped kernel poisons its CUDA context) and
# keeps the logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
T="timeout 600 python -m pytest tests/test_gpu_kernels.py -q --timeout 300 -m gpu -p no:cacheprovider"
$T -k "not tc_gemm and not deepsea and not forward and not fused" > gpurun_out/t_hbm.log 2>&1; echo "hbm-kernels exit $?"
$T -k "tc_gemm" > gpurun_out/t_tc.log 2>&1; echo "tc-gemm exit $?"
SB_TC_PER_TAP=0 $T -k "deepsea or forward or fused" > gpurun_out/t_net.log 2>&1; echo "net exit $?"
SB_TC_PER_TAP=1 $T -k "deepsea_forward" > gpurun_out/t_net_pertap.log 2>&1; echo "net per-tap exit $?"
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_analyze.py -q --timeout 300 -m gpu -p no:cacheprovider > gpurun_out/t_golden.log 2>&1; echo "golden exit $?"
timeout 900 python -m pytest tests/test_gpu_train.py -q --timeout 600 -m gpu -p no:cacheprovider > gpurun_out/t_train.log 2>&1; echo "train exit $?"
tail -n 25 gpurun_out/t_train.log gpurun_out/t_golden.log gpurun_out/t_hbm.log gpurun_out/t_tc.log gpurun_out/t_net.log gpurun_out/t_net_pertap.log
