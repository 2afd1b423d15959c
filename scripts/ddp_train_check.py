"""2-GPU check of data-parallel training (run under torchrun): every rank steps on its own half batch with
one NCCL all-reduce of the flat gradient vector; rank 0 also steps a second trainer on the full batch and
the two must agree (mean-BCE gradients average exactly)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import sb_models as M
from selene_b200.train_model import B200Trainer

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
torch.manual_seed(0)
model = M.DeepSEA(1000, 919)
rng = np.random.default_rng(1)
n = 8 * world
codes = rng.integers(0, 4, (n, 1000))
x = np.zeros((n, 1000, 4), np.float32)
for b in range(4):
    x[codes == b, b] = 1
y = (rng.random((n, 919)) < 0.1).astype(np.float32)
lo, hi = rank * 8, rank * 8 + 8
tr = B200Trainer(model, 1000, lr=0.05, momentum=0.9, weight_decay=1e-6)
xs, ys = torch.from_numpy(x[lo:hi]).cuda(), torch.from_numpy(y[lo:hi]).cuda()
losses = [tr.step(xs, ys) for _ in range(3)]
w_dp, _ = tr.layer_params(1)
ok = True
if rank == 0:
    dist.barrier()
    ref = B200Trainer(model, 1000, lr=0.05, momentum=0.9, weight_decay=1e-6)
    saved = dist.group.WORLD
    xf, yf = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    for _ in range(3):
        ref.forward_backward(xf, yf)
        ref.sgd_step()
    w_ref, _ = ref.layer_params(1)
    err = np.abs(w_dp - w_ref).max()
    ok = err <= 1e-5 * max(1.0, np.abs(w_ref).max())
    print("ddp check: world %d, max |w_dp - w_full_batch| = %.3g, losses %s -> %s" % (world, err, losses, "OK" if ok else "MISMATCH"))
else:
    dist.barrier()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
