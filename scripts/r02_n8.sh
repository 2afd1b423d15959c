#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err
echo "bench ref n=$N exit $?"; cut -c1-200 gpurun_out/bench_ref_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench n=$N exit $?"
python -c "
import json,sys; d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1]); print(d['value'], d['e2e']['value'], d['ms_per_step']); print('train', {k:v for k,v in d['train'].items() if k not in ('workload',)}); print('strong', {k:v for k,v in d['strong'].items() if k!='what'})"
tail -3 gpurun_out/bench_n$N.err
