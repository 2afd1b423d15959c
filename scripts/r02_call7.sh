#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/score_parity_stats.txt
timeout 1700 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "all gpu tests exit $?"
tail -4 gpurun_out/t_all.log
bash scripts/mutation_check.sh run; echo "mutation check exit $?"
python bench.py --steps 6 --warmup 3 --no-cpu --strong-variants 0 2> gpurun_out/bench.err | grep '^{' > gpurun_out/bench.json; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'])
for r in d['roofline_hbm']: print('%-44s %7.1f GB/s  frac %.3f  %.3f ms' % (r['kernel'], r['achieved'], r['frac'], r['ms']))"
SMALL="python bench.py --steps 2 --warmup 1 --no-cpu --no-hbm --strong-variants 0"
$SMALL > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:tc_layer -s 5 -c 5 -o gpurun_out/prof_tc -f $SMALL > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
python scripts/prof_hbm.py > gpurun_out/prof_hbm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'window_kernel|label_bins|ism_expand' -c 16 -o gpurun_out/prof_hbm -f python scripts/prof_hbm.py > gpurun_out/ncu_hbm.log 2>&1
echo "ncu hbm exit $?"
