#!/bin/bash
# 2-GPU call: the byte-identical sharded / gathered VEP file test + bench at N=2 (strong + train legs under NCCL)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_analyze.py -x -q -m gpu -p no:cacheprovider -k two_processes > gpurun_out/t_2gpu.log 2>&1; echo "2-gpu vep test exit $?"
tail -3 gpurun_out/t_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n=2 exit $?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_n2.json') if l.startswith('{')][-1]); print(d['value'], d['e2e']['value']); print('train', d.get('train')); print('strong', {k:v for k,v in d['strong'].items() if k!='what'})"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
echo "bench ref n=2 exit $?"; cut -c1-300 gpurun_out/bench_ref_n2.json
