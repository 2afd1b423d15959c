#!/bin/bash
# 2-GPU: bench (all legs incl. sampler-fed training and bf16 gradient all-reduce) + DDP equality check (fp32 path)
mkdir -p gpurun_out
N=2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench n=$N exit $?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_n2.json') if l.startswith('{')][-1]); print(d['value'], d['e2e']['value']); print('train', d.get('train')); print('strong', {k:v for k,v in d['strong'].items() if k!='what'})"
tail -3 gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  scripts/ddp_train_check.py > gpurun_out/ddp_n$N.log 2>&1
echo "ddp n=$N exit $?"; grep "ddp check" gpurun_out/ddp_n$N.log || tail -5 gpurun_out/ddp_n$N.log
python bench.py --steps 4 --warmup 3 --no-cpu --no-hbm --strong-variants 0 2> gpurun_out/bench1.err | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('n=1 train', d['train'])"
