#!/bin/bash
mkdir -p gpurun_out
python bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; echo "configs exit $?"
python -c "
import json
for l in open('gpurun_out/configs.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print(json.dumps({k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items()})[:260])"
tail -3 gpurun_out/configs.err
