import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import sb_models as M
from selene_b200.nets import B200Net, layers_from_module
from selene_b200 import _lib
torch.manual_seed(0)
net = B200Net(layers_from_module(M.DanQ(1000, 919).eval()), 1000)
lib = _lib.load()
for n in (2048, 3000, 4096):
    codes = torch.from_numpy(np.random.default_rng(0).integers(0, 4, (n, 1000)).astype(np.uint8)).cuda()
    net.set_recurrence(64, 1)
    for _ in range(2):
        net.forward_codes(codes)
    torch.cuda.synchronize()
    lib.sb_net_set_timing(net._h, 1)
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    net.forward_codes(codes)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    import ctypes as C
    ms = []
    for li in range(lib.sb_net_n_layers(net._h)):
        tm, ln = C.c_double(), C.c_int()
        lib.sb_net_layer_time(net._h, li, C.byref(tm), C.byref(ln))
        ms.append(round(tm.value, 3))
    lib.sb_net_set_timing(net._h, 0)
    print("n=%d: host issue %.2f ms, wall %.2f ms, device %.2f ms, per-layer ms %s" % (n, (t1 - t0) * 1e3, (t2 - t0) * 1e3, e0.elapsed_time(e1), ms))
