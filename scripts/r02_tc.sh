#!/bin/bash
# tensor path: parity tests of the conv stacks + score epilogue, then the bench line (layer times)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py tests/test_gpu_score_parity.py tests/test_gpu_analyze.py -x -q -m gpu -p no:cacheprovider > gpurun_out/t_tc.log 2>&1; echo "tc tests exit $?"
tail -3 gpurun_out/t_tc.log
for v in "$@"; do
  echo "== $v"
  env $v python bench.py --steps 6 --warmup 3 --no-cpu --no-hbm --no-train --strong-variants 0 2> gpurun_out/bench_tc.err | grep '^{' > gpurun_out/bench_tc.json; echo "bench exit $?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_tc.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'])
print(d['roofline']['layer_ms_per_step'], d['roofline']['whole_net_frac'], d['clocks'])"
done
