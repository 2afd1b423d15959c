#!/bin/bash
# round-2 evidence: bench line, ncu launch list of a small bench step, ncu --set full of the tensor kernels and of the
# HBM-bound kernels (each ncu pass only after the same command exited 0 without ncu)
mkdir -p gpurun_out
python bench.py --steps 8 --warmup 3 2> gpurun_out/bench.err | grep '^{' > gpurun_out/bench.json; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'], d['ms_per_step']); print(d['roofline']['layer_ms_per_step'], d['roofline']['whole_net_frac']); print('train', d['train']['ms_per_step']); print('cpu', d['cpu_baseline']['value'])"
SMALL="python bench.py --steps 2 --warmup 1 --no-cpu --no-hbm --no-train --strong-variants 0"
$SMALL > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:'tc_layer|conv1_mma' -s 5 -c 12 -o gpurun_out/prof_tc -f $SMALL > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
python scripts/prof_hbm.py --ncu > gpurun_out/prof_hbm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'window_kernel|label_bins|ism_expand' -c 16 -o gpurun_out/prof_hbm -f python scripts/prof_hbm.py --ncu > gpurun_out/ncu_hbm.log 2>&1
echo "ncu hbm exit $?"
