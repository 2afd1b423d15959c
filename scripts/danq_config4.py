"""BASELINE.json configs[3]: DanQ (conv + BiLSTM, 1000 bp, 919 targets) in-silico mutagenesis of 10 000 synthetic
sequences (3 000 single-base mutants each = 3e7 mutant predictions), abs_diffs + logits, through
AnalyzeSequences.ism_scores, two sequences in flight on two CUDA streams.  Prints one JSON line.
    python scripts/danq_config4.py [n_sequences]
"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import sb_models as M
from selene_b200.predict.model_predict import AnalyzeSequences
from selene_b200.sequences import Genome

n_seq = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
n_streams = int(sys.argv[2]) if len(sys.argv) > 2 else 2
feats = ["f%d" % i for i in range(919)]
Genome.update_bases_order(['A', 'C', 'G', 'T'])
torch.manual_seed(0)
models = [M.DanQ(1000, 919)]
for _ in range(n_streams - 1):
    models.append(M.DanQ(1000, 919))
    models[-1].load_state_dict(models[0].state_dict())
azs = [AnalyzeSequences(m, sequence_length=1000, features=feats, reference_sequence=Genome) for m in models]
streams = [torch.cuda.Stream() for _ in range(n_streams)]
rng = np.random.default_rng(5)
letters = np.array(list("ACGT"))
seqs = ["".join(letters[rng.integers(0, 4, 1000)].tolist()) for _ in range(min(n_seq, 512))]   # cycled; every one distinct
res = {}
for to_host, n in ((False, n_seq), (True, min(n_seq, 400))):
    for k in range(2 * n_streams):
        with torch.cuda.stream(streams[k % n_streams]):
            azs[k % n_streams].ism_scores(seqs[k], ("abs_diffs", "logits"), to_host=to_host)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(n):
        with torch.cuda.stream(streams[k % n_streams]):
            azs[k % n_streams].ism_scores(seqs[k % len(seqs)], ("abs_diffs", "logits"), to_host=to_host)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res["to_host" if to_host else "device_resident"] = {"sequences": n, "seconds": dt, "mutants_per_s": 3000 * n / dt,
                                                        "ms_per_sequence": dt / n * 1e3}
fl = azs[0].net.flops_per_seq
res["config"] = "DanQ ISM, %d x 1000-bp sequences x 3000 mutants, batch-axis recurrence in groups of 64, bf16, %d streams" % (n_seq, n_streams)
res["flops_per_mutant"] = fl
res["tflops_device_resident"] = res["device_resident"]["mutants_per_s"] * fl / 1e12
print(json.dumps(res))
