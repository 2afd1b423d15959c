#!/bin/bash
# usage: r02_ncu_one.sh <kernel regex> <out name> [skip] [count]  — one ncu --set full capture over the small bench
mkdir -p gpurun_out
SMALL="python bench.py --steps 2 --warmup 1 --no-cpu --no-hbm --no-train --strong-variants 0"
$SMALL > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$1 -s ${3:-2} -c ${4:-2} -o gpurun_out/$2 -f $SMALL > gpurun_out/ncu_$2.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_$2.log
