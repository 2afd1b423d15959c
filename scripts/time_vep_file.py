"""File-to-file VEP with the stage timings of AnalyzeSequences.last_timing: VCF -> abs_diffs + logits TSV.
usage: python scripts/time_vep_file.py [n_variants] [block ...]   (SB_WRITER_MMAP / SB_WRITER_THREADS from the environment)"""
import json, os, shutil, sys, tempfile, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench as B
import sb_models as M
from selene_b200.sequences import Genome
from selene_b200.predict.model_predict import AnalyzeSequences

nv = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
blocks = [int(a) for a in sys.argv[2:]] or [32768]
names, seqs = B.synth_genome_arrays(240.0)
G = Genome.from_sequences(names, seqs)
Genome.update_bases_order(['A', 'C', 'G', 'T'])
per = len(seqs[0])
feats = ["f%d" % i for i in range(919)]
tmpd = tempfile.mkdtemp(prefix="sb_vep_")
vcf = os.path.join(tmpd, "v.vcf")
vr = np.random.default_rng(2024)
ci_, pos_, alt_ = vr.integers(0, len(names), nv), vr.integers(600, per - 600, nv), vr.integers(0, 4, nv)
with open(vcf, "w") as fh:
    fh.write("#CHROM\tPOS\tID\tREF\tALT\n")
    for i in range(nv):
        ref_ = chr(seqs[ci_[i]][pos_[i] - 1]).upper()
        fh.write("%s\t%d\trs%d\t%s\t%s\n" % (names[ci_[i]], pos_[i], i, ref_ if ref_ in "ACGT" else "A", "ACGT"[alt_[i]]))
torch.manual_seed(0)
az = AnalyzeSequences(M.DeepSEA(1000, 919), sequence_length=1000, features=feats, reference_sequence=G)
ref_sizes = None
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    az.variant_effect_prediction(vcf, ["abs_diffs", "logits"], output_dir=os.path.join(tmpd, "w"))
    for blk in blocks:
        az.vep_block_variants = blk
        best = None
        for rep in range(2):
            out = os.path.join(tmpd, "o%d_%d" % (blk, rep))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            az.variant_effect_prediction(vcf, ["abs_diffs", "logits"], output_dir=out)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            sizes = sorted((f, os.path.getsize(os.path.join(out, f))) for f in os.listdir(out))
            if ref_sizes is None:
                ref_sizes = sizes
                import hashlib
                ref_md5 = [hashlib.md5(open(os.path.join(out, f), "rb").read()).hexdigest() for f, _ in sizes]
            else:
                import hashlib
                md5 = [hashlib.md5(open(os.path.join(out, f), "rb").read()).hexdigest() for f, _ in sizes]
                assert sizes == ref_sizes and md5 == ref_md5, "output files differ between variants"
            shutil.rmtree(out)
            if best is None or dt < best[0]:
                best = (dt, dict(az.last_timing))
        print(json.dumps({"block": blk, "mmap": os.environ.get("SB_WRITER_MMAP", "0"), "threads": os.environ.get("SB_WRITER_THREADS", ""),
                          "variants": nv, "seconds": round(best[0], 4), "variants_per_s": round(nv / best[0]),
                          "timing": {k: round(v, 4) for k, v in best[1].items()}}), flush=True)
shutil.rmtree(tmpd, ignore_errors=True)
