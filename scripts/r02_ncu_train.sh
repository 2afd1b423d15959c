#!/bin/bash
# usage: r02_ncu_train.sh <kernel regex> <out name> [skip] [count] — one ncu --set full capture over scripts/prof_train.py
mkdir -p gpurun_out
python scripts/prof_train.py > gpurun_out/train_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/train_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$1 -s ${3:-2} -c ${4:-1} -o gpurun_out/$2 -f python scripts/prof_train.py > gpurun_out/ncu_$2.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_$2.log
