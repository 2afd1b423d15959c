import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import sb_models as M
from selene_b200.train_model import B200Trainer
torch.manual_seed(0)
mdl = M.DeepSEA(1000, 919)
tr = B200Trainer(mdl, 1000, lr=0.01, momentum=0.9, weight_decay=1e-6, dropout=B200Trainer.dropout_of_module(mdl), precision="bf16")
rng = np.random.default_rng(0)
codes = rng.integers(0, 4, (64, 1000))
x = np.zeros((64, 1000, 4), np.float32)
for b in range(4): x[codes == b, b] = 1
y = (rng.random((64, 919)) < 0.1).astype(np.float32)
xb, yb = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
for _ in range(3):
    tr.step(xb, yb)
torch.cuda.synchronize()
print("ok")
