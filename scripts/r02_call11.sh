#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "all gpu tests exit $?"
tail -4 gpurun_out/t_all.log
bash scripts/r02_tcb.sh "SB_X=1" "SB_CHUNK=2048" "SB_CHUNK=8192" "SB_CONV1_AHEAD=0 SB_CONV1_MMA=0 SB_TC_IM2COL=0 SB_CHUNK=8192"
