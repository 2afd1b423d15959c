#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "all gpu tests exit $?"
tail -4 gpurun_out/t_all.log
python bench.py --steps 6 --warmup 3 --strong-variants 262144 2> gpurun_out/bench.err | grep '^{' > gpurun_out/bench.json; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value']); print('train', d.get('train')); print('strong', d.get('strong')); print('cpu', d.get('cpu_baseline'))
print(d['roofline']['layer_ms_per_step'], d['roofline']['whole_net_frac'])
for r in d['roofline_hbm']: print('%-44s %7.1f GB/s  frac %.3f  %.3f ms' % (r['kernel'], r['achieved'], r['frac'], r['ms']))"
tail -3 gpurun_out/bench.err
