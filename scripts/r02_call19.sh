#!/bin/bash
# 2-GPU: bench without the CPU / HBM / strong legs (VEP + data-parallel training with the bf16 gradient all-reduce)
mkdir -p gpurun_out
N=2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 6 --warmup 3 --no-cpu --no-hbm --strong-variants 0 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench n=$N exit $?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_n2.json') if l.startswith('{')][-1]); print(d['value'], d['e2e']['value']); print('train', d.get('train'))"
tail -3 gpurun_out/bench_n$N.err
