#!/bin/bash
# HBM-bound kernels: parity tests of stage 1 / 2a + their roofline lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py tests/test_samples.py -x -q -m gpu -p no:cacheprovider -k "encode or labels or snv or ism or golden or genome or sampler or samples or unpack or pack" > gpurun_out/t_hbm.log 2>&1; echo "hbm tests exit $?"
tail -3 gpurun_out/t_hbm.log
python scripts/prof_hbm.py > gpurun_out/hbm_lines.jsonl 2> gpurun_out/hbm_lines.err; echo "hbm exit $?"
python -c "
import json
for l in open('gpurun_out/hbm_lines.jsonl'):
    if l.startswith('{'):
        r=json.loads(l); print('%-44s %7.1f GB/s  frac %.3f  %.3f ms' % (r['kernel'], r['achieved'], r['frac'], r['ms']))"
