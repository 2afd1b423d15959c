#!/bin/bash
# bench-only variants: each argument is an env assignment list
mkdir -p gpurun_out
for v in "$@"; do
  echo "== $v"
  env $v python bench.py --steps 8 --warmup 3 --no-cpu --no-hbm --no-train --strong-variants 0 2> gpurun_out/bench_tc.err | grep '^{' > gpurun_out/bench_tc.json; echo "bench exit $?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_tc.json')); print('value %.0f e2e %.0f ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
print({k: round(v,2) for k,v in d['roofline']['layer_ms_per_step'].items()}, round(d['roofline']['whole_net_frac'],3), d['clocks']['sm_mhz'])"
done
