#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "all gpu tests exit $?"
tail -4 gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
( time python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_default.json') if l.startswith('{')][-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['whole_net_frac'])"
