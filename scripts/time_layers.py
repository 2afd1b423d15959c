"""Per-layer CUDA-event times and FLOP rates of one of the reference's conv stacks.
Usage: python scripts/time_layers.py heartenn|seqweaver|deeperdeepsea|deepsea [rows]"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import sb_models as M
from selene_b200.nets import B200Net, layers_from_module
from selene_b200 import _lib
arch = sys.argv[1] if len(sys.argv) > 1 else "heartenn"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
torch.manual_seed(0)
model = {"heartenn": lambda: M.HeartENN(1000, 90), "seqweaver": lambda: M.Seqweaver(217),
         "deeperdeepsea": lambda: M.DeeperDeepSEA(1000, 919), "deepsea": lambda: M.DeepSEA(1000, 919)}[arch]().eval()
net = B200Net(layers_from_module(model), 1000)
lib = _lib.load()
lib.sb_net_layer_flops_per_seq.restype = C.c_double
codes = torch.from_numpy(np.random.default_rng(0).integers(0, 4, (n, 1000)).astype(np.uint8)).cuda()
for _ in range(2):
    net.forward_codes(codes)
torch.cuda.synchronize()
lib.sb_net_set_timing(net._h, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
net.forward_codes(codes)
e1.record()
torch.cuda.synchronize()
tot = e0.elapsed_time(e1)
print("%s, %d rows: %.3f ms, %.1f TFLOP/s" % (arch, n, tot, n * net.flops_per_seq / tot / 1e9))
for li in range(lib.sb_net_n_layers(net._h)):
    tm, ln = C.c_double(), C.c_int()
    lib.sb_net_layer_time(net._h, li, C.byref(tm), C.byref(ln))
    fl = lib.sb_net_layer_flops_per_seq(net._h, li)
    print("  layer %d: %.3f ms (%d launches), %.1f MFLOP/seq, %.1f TFLOP/s" % (li, tm.value, ln.value, fl / 1e6,
                                                                              n * fl / max(tm.value, 1e-9) / 1e9))
