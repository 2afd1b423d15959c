"""Config 3 on N GPUs (run under torchrun): DeepSEA training step, bf16 tensor-core path, batch 64 per GPU,
one NCCL all-reduce of the flat fp32 gradient vector (52.5 M floats) per step.  Rank 0 prints one JSON line with
the aggregate samples/s (device time, max over ranks)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import sb_models as M
from selene_b200.train_model import B200Trainer

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = 30
torch.manual_seed(0)
model = M.DeepSEA(1000, 919)
tr = B200Trainer(model, 1000, lr=0.01, momentum=0.9, weight_decay=1e-6, dropout=B200Trainer.dropout_of_module(model),
                 precision="bf16")
rng = np.random.default_rng(1 + rank)
codes = rng.integers(0, 4, (batch, 1000))
x = np.zeros((batch, 1000, 4), np.float32)
for b in range(4):
    x[codes == b, b] = 1
y = (rng.random((batch, 919)) < 0.1).astype(np.float32)
xs, ys = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
for _ in range(5):
    tr.step(xs, ys)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
overlap = os.environ.get("SB_DDP_OVERLAP", "1") == "1"
if world > 1 and overlap:
    tr._enable_overlap()
for _ in range(steps):
    tr.forward_backward(xs, ys)
    tr.allreduce_grads(overlap=overlap)
    tr.sgd_step()
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t.item()) / steps
    print(json.dumps({"config": "3: DeepSEA training step, data parallel", "n_gpus": world, "batch_per_gpu": batch,
                      "precision": "bf16", "ms_per_step": ms, "samples_per_s": world * batch / ms * 1e3,
                      "allreduce_bytes_per_step": int(tr.grads.numel()) * 4,
                      "overlap_tail_allreduce": bool(world > 1 and overlap)}))
if world > 1:
    dist.destroy_process_group()
