#!/bin/bash
# N-GPU bench line (run with: gpurun --gpus N -- bash scripts/gpu_multi.sh N)
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench n=$N exit $?"; cat gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
