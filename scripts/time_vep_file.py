"""End-to-end variant_effect_prediction from a VCF file to the TSV files (the call a selene user makes)."""
import cProfile, os, pstats, sys, tempfile, time
import numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import sb_models as M
import bench as B
from selene_b200.sequences import Genome
from selene_b200.predict.model_predict import AnalyzeSequences
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
names, seqs = B.synth_genome_arrays(24.0)
G = Genome.from_sequences(names, seqs)
Genome.update_bases_order(['A', 'C', 'G', 'T'])
torch.manual_seed(0)
az = AnalyzeSequences(M.DeepSEA(1000, 919), sequence_length=1000, features=["f%d" % i for i in range(919)], reference_sequence=G)
rng = np.random.default_rng(2024)
per = len(seqs[0])
tmp = tempfile.mkdtemp(prefix="sb_vep_")
vcf = os.path.join(tmp, "v.vcf")
with open(vcf, "w") as fh:
    fh.write("#CHROM\tPOS\tID\tREF\tALT\n")
    ci = rng.integers(0, len(names), n); pos = rng.integers(600, per - 600, n); alt = rng.integers(0, 4, n)
    for i in range(n):
        ref = chr(seqs[ci[i]][pos[i] - 1]).upper()
        ref = ref if ref in "ACGT" else "A"
        fh.write("%s\t%d\trs%d\t%s\t%s\n" % (names[ci[i]], pos[i], i, ref, "ACGT"[alt[i]]))
for rep in range(2):
    out = os.path.join(tmp, "o%d" % rep)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if rep == 1 and len(sys.argv) > 2:
        pr = cProfile.Profile(); pr.enable()
    az.variant_effect_prediction(vcf, ["abs_diffs", "logits"], output_dir=out)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if rep == 1 and len(sys.argv) > 2:
        pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
    sz = sum(os.path.getsize(os.path.join(out, f)) for f in os.listdir(out))
    print("stages:", {k: round(v, 3) for k, v in az.last_timing.items()})
    print("rep %d: %d variants in %.2f s = %.0f variants/s (%.0f predictions/s), %.1f MB written" % (rep, n, dt, n / dt, 2 * n / dt, sz / 1e6))
