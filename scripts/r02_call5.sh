#!/bin/bash
# new window / SNV / label kernels + writers + VCF table: whole suite, bench, file-to-file VEP, ncu of the HBM kernels
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "all gpu tests exit $?"
tail -12 gpurun_out/t_all.log
python bench.py --steps 6 --warmup 3 --no-cpu 2> gpurun_out/bench.err | grep '^{' > gpurun_out/bench.json; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'])
for r in d['roofline_hbm']: print('%-30s %7.1f GB/s  frac %.3f  %.3f ms' % (r['kernel'], r['achieved'], r['frac'], r['ms']))"
python scripts/time_vep_file.py 100000 > gpurun_out/vep_file.log 2>&1; grep -E "stages|rep" gpurun_out/vep_file.log
python scripts/prof_hbm.py > gpurun_out/prof_hbm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'window_kernel|label_bins|ism_expand' -c 16 -o gpurun_out/prof_hbm -f python scripts/prof_hbm.py > gpurun_out/ncu_hbm.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_hbm.log
