import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from selene_b200 import _lib
lib = _lib.load()
F, nrows, nchrom, per = 919, 400000, 24, 2000000
lr = np.random.default_rng(7)
c_i = lr.integers(0, nchrom, nrows).astype(np.int32)
s_i = lr.integers(0, per - 3000, nrows).astype(np.int32)
ln = np.clip(lr.lognormal(np.log(300), 0.8, nrows), 50, 2000).astype(np.int32)
f_i = lr.integers(0, F, nrows).astype(np.int32)
h = C.c_void_p()
e_i = np.ascontiguousarray(s_i + ln)
_lib.check(lib.sb_intervals_create(_lib.ptr(c_i), _lib.ptr(s_i), _lib.ptr(e_i), _lib.ptr(f_i), nrows, F, nchrom, 0, C.byref(h)), "create")
nq = 1 << 19
qc = torch.from_numpy(lr.integers(0, nchrom, nq).astype(np.int32)).cuda()
qs_np = lr.integers(0, per - 200, nq).astype(np.int32)
qs = torch.from_numpy(qs_np).cuda(); qe = torch.from_numpy(qs_np + 200).cuda()
thr = torch.full((F,), 99, dtype=torch.int32, device="cuda")
o = torch.empty((nq, F), dtype=torch.int64, device="cuda")
for _ in range(4):
    _lib.check(lib.sb_label_bins(h, _lib.ptr(qc), _lib.ptr(qs), _lib.ptr(qe), nq, _lib.ptr(thr), _lib.SB_I64, _lib.ptr(o), _lib.stream_ptr()), "label")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    _lib.check(lib.sb_label_bins(h, _lib.ptr(qc), _lib.ptr(qs), _lib.ptr(qe), nq, _lib.ptr(thr), _lib.SB_I64, _lib.ptr(o), _lib.stream_ptr()), "label")
e1.record(); torch.cuda.synchronize()
print("ms per launch", e0.elapsed_time(e1) / 10, "positives", o.float().mean().item())
