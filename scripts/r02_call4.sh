#!/bin/bash
# 2 GPUs: the whole -m gpu suite (incl. the 2-rank VEP file identity test), then the N=2 bench with the warmed gather
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "all gpu tests exit $?"
tail -15 gpurun_out/t_all.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 2> gpurun_out/bench_n2.err | grep '^{' > gpurun_out/bench_n2.json; echo "bench n2 exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print(d['value'], d['e2e']['value']); s=d['strong']; s.pop('what'); print(s)"
python scripts/time_vep_file.py 100000 > gpurun_out/vep_file.log 2>&1; tail -5 gpurun_out/vep_file.log
