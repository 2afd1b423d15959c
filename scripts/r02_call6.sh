#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py tests/test_gpu_train.py -x -q -m gpu -p no:cacheprovider -k "encode or snv or label or golden or genome or bases or train or sampler" > gpurun_out/t_k.log 2>&1; echo "kernel tests exit $?"
tail -4 gpurun_out/t_k.log
python bench.py --steps 6 --warmup 3 --no-cpu --strong-variants 0 2> gpurun_out/bench.err | grep '^{' > gpurun_out/bench.json; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'])
for r in d['roofline_hbm']: print('%-30s %7.1f GB/s  frac %.3f  %.3f ms' % (r['kernel'], r['achieved'], r['frac'], r['ms']))"
python scripts/prof_hbm.py > gpurun_out/prof_hbm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'window_kernel|label_bins|ism_expand' -c 16 -o gpurun_out/prof_hbm -f python scripts/prof_hbm.py > gpurun_out/ncu_hbm.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_hbm.log
