"""Helper of test_vep_two_processes_write_identical_files: runs AnalyzeSequences.variant_effect_prediction on a
seeded VCF (SNVs on both strands, multi-alt rows, indels, NA rows) as a single process or as a torchrun rank.

    python multi_gpu_vep_check.py <workdir> single|shards|gather
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, HERE)


def main():
    import torch
    import torch.distributed as dist
    import sb_helpers as H
    import sb_models as M
    work, mode = sys.argv[1], sys.argv[2]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from selene_b200.predict.model_predict import AnalyzeSequences
    from selene_b200.sequences import Genome
    g = H.synth_genome(n_chrom=3, length=30000)
    fa = os.path.join(work, "g_%s_%d.fa" % (mode, rank))
    H.write_fasta(g, fa)
    names = sorted(g.keys())
    vcf = os.path.join(work, "v.vcf")
    if rank == 0 and not os.path.exists(vcf):
        rng = np.random.default_rng(3)
        with open(vcf + ".tmp", "w") as fh:
            fh.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tSTRAND\n")
            for i in range(4001):
                c = names[rng.integers(0, 3)]
                pos = int(rng.integers(1, len(g[c])))
                ref = g[c][pos - 1].upper()
                ref = ref if ref in "ACGT" else "A"
                kind = rng.random()
                if kind < 0.9:
                    alt = "ACGT"[rng.integers(0, 4)]
                elif kind < 0.94:
                    alt = "A,T"
                elif kind < 0.97:
                    alt = ref + "ACG"[:rng.integers(1, 4)]
                else:
                    ref, alt = g[c][pos - 1:pos + 2].upper().replace("N", "A"), "-"
                fh.write("%s\t%d\tv%d\t%s\t%s\t%s\n" % (c, pos, i, ref, alt, "+-."[rng.integers(0, 3)]))
        os.replace(vcf + ".tmp", vcf)
    if world > 1:
        dist.barrier()
    Genome.update_bases_order(['A', 'C', 'G', 'T'])
    torch.manual_seed(0)
    az = AnalyzeSequences(M.DeepSEA(1000, 919), sequence_length=1000, features=["f%d" % i for i in range(919)],
                          reference_sequence=Genome(fa), precision="bf16", device=local_rank)
    if mode != "single":
        az.multi_gpu_output = mode
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        az.variant_effect_prediction(vcf, ["abs_diffs", "logits", "predictions"], output_dir=os.path.join(work, mode),
                                     strand_index=5)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
