#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -x -q -m gpu -p no:cacheprovider > gpurun_out/t_train.log 2>&1; echo "train tests exit $?"
tail -5 gpurun_out/t_train.log
for v in "$@"; do
echo "== $v"
env $v python - <<'PY'
import bench, json, sys
sys.path.insert(0, 'tests')
from sb_models import DeepSEA
import torch
torch.cuda.set_device(0)
r = bench.train_leg(DeepSEA, 0, 1)
print(json.dumps({k: r[k] for k in ('ms_per_step', 'samples_per_s', 'loss', 'tflops')}))
PY
done
