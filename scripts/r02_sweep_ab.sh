#!/bin/bash
mkdir -p gpurun_out
for v in "SB_X=1" "SB_TC_IM2COL=0" "SB_X=2" ; do
echo "== $v"
env $v python bench_configs.py --only sweep 2> gpurun_out/configs_sweep.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        if d.get('variants_per_batch') in (1024, 4096): print('%-60s %5d %8.3f ms %10.0f pred/s %.3f' % (d['config'][:60], d['variants_per_batch'], d['ms'], d['predictions_per_s'], d['tensor_frac_of_sustained']))"
done
