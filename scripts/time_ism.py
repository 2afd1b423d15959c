import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import sb_models as M
from selene_b200.sequences import Genome
from selene_b200.predict.model_predict import AnalyzeSequences
from selene_b200.predict._in_silico_mutagenesis import ism_mutations_device
from selene_b200 import _lib
torch.manual_seed(0)
which = sys.argv[1] if len(sys.argv) > 1 else "danq"
model = M.DanQ(1000, 919) if which == "danq" else M.DeepSEA(1000, 919)
az = AnalyzeSequences(model, sequence_length=1000, features=["f%d" % i for i in range(919)], reference_sequence=Genome)
r = np.random.default_rng(0)
seq = "".join(np.array(list("ACGT"))[r.integers(0, 4, 1000)].tolist())
for _ in range(3):
    az.ism_scores(seq, ("abs_diffs", "logits"), to_host=False)
torch.cuda.synchronize()
def T(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
    print("%-28s %.2f ms" % (label, (time.perf_counter() - t0) * 1e3)); return out
pos, alt, codes = T("ism_mutations_device", lambda: ism_mutations_device(seq, 0, 1000, Genome))
az.net.set_recurrence(64, 1)
base = T("base forward (n=1)", lambda: az.net.forward_codes(codes, precision="bf16"))
m = int(pos.numel())
src = torch.zeros(m, dtype=torch.int32, device="cuda")
T("forward_score (n=%d)" % m, lambda: az.net.forward_score(codes, src, pos, alt, m, _lib.SB_PAIR_BASELINE, precision="bf16", base_pred=base, want=("abs_diffs", "logits")))
T("forward_codes (n=%d)" % m, lambda: az.net.forward_codes(codes, src, pos, alt, n=m, precision="bf16"))
T("whole ism_scores", lambda: az.ism_scores(seq, ("abs_diffs", "logits"), to_host=False))
T("whole ism_scores to host", lambda: az.ism_scores(seq, ("abs_diffs", "logits"), to_host=True))
