#!/bin/bash
# N-GPU runs (gpurun --gpus N -- bash scripts/gpu_multi.sh N): scaling bench line + DDP training check
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench n=$N exit $?"; cut -c1-400 gpurun_out/bench_n$N.json; tail -2 gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  scripts/ddp_train_check.py > gpurun_out/ddp_n$N.log 2>&1
echo "ddp n=$N exit $?"; grep "ddp check" gpurun_out/ddp_n$N.log || tail -5 gpurun_out/ddp_n$N.log
