import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import sb_models as M
from selene_b200.sequences import Genome
from selene_b200.predict.model_predict import AnalyzeSequences
torch.manual_seed(0)
az = AnalyzeSequences(M.DanQ(1000, 919), sequence_length=1000, features=["f%d" % i for i in range(919)], reference_sequence=Genome)
r = np.random.default_rng(0)
seq = "".join(np.array(list("ACGT"))[r.integers(0, 4, 1000)].tolist())
az.ism_scores(seq, ("abs_diffs",), to_host=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
az.ism_scores(seq, ("abs_diffs", "logits"), to_host=False)
torch.cuda.synchronize()
print("one ISM sequence: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
