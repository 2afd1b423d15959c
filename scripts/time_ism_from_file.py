import os, sys, time, tempfile, numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import sb_models as M
from selene_b200.sequences import Genome
from selene_b200.predict.model_predict import AnalyzeSequences
torch.manual_seed(0)
which = sys.argv[1]
model = M.DanQ(1000, 919) if which == "danq" else M.DeepSEA(1000, 919)
az = AnalyzeSequences(model, sequence_length=1000, features=["f%d" % i for i in range(919)], reference_sequence=Genome)
r = np.random.default_rng(11)
tmp = tempfile.mkdtemp(prefix="ismf_")
fa = os.path.join(tmp, "q.fa")
n = 24
with open(fa, "w") as fh:
    for i in range(n):
        fh.write(">seq%d\n%s\n" % (i, "".join(np.array(list("ACGT"))[r.integers(0, 4, 1000)].tolist())))
az.in_silico_mutagenesis_from_file(fa, ["abs_diffs", "logits"], os.path.join(tmp, "w"))
torch.cuda.synchronize(); t0 = time.perf_counter()
az.in_silico_mutagenesis_from_file(fa, ["abs_diffs", "logits"], os.path.join(tmp, "o"))
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("%s: %d sequences in %.2f s = %.1f ms/sequence = %.0f mutants/s incl. files" % (which, n, dt, dt / n * 1e3, 3000 * n / dt))
