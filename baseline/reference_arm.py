"""Drives the UNMODIFIED reference (baseline/_ref, installed by baseline/build_ref.py) through its own public
API for bench.py's CPU legs: ``selene_sdk.predict.AnalyzeSequences.variant_effect_prediction`` on a FASTA file
and a VCF file, DeepSEA from the reference's own ``models/deepsea.py``, ``use_cuda=False``.

Nothing of selene_b200 is on this path.  What this module adds around the stock call:
  * the four third-party imports missing from the image are satisfied by baseline/shims/ (I/O only);
  * ``write=False`` replaces ``predict_handlers.handler.write_to_tsv_file`` — the function the stock handlers
    call to append their rows to the TSV — by an in-memory collector for the duration of the call, so the number
    compares like with like (our e2e leg returns score matrices, it does not write TSV).  The handlers,
    the per-variant loop, the encoders, ``predict`` and the model run untouched.  ``write=True`` times the stock
    call including its files.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
SHIMS = os.path.join(HERE, "shims")


def available():
    return os.path.isdir(os.path.join(REF_DIR, "selene_sdk")) and os.path.isdir(os.path.join(REF_DIR, "selene_models"))


def _import_reference():
    for p in (SHIMS, REF_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    import selene_sdk  # noqa: F401
    from selene_sdk.sequences import Genome
    from selene_sdk.predict import AnalyzeSequences
    from selene_sdk.predict.predict_handlers import handler as handler_mod
    spec = importlib.util.spec_from_file_location("selene_models_deepsea", os.path.join(REF_DIR, "selene_models", "deepsea.py"))
    deepsea = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(deepsea)
    return Genome, AnalyzeSequences, handler_mod, deepsea


def write_fasta(names, seqs, path, width=60):
    """names: list of str; seqs: list of uint8 ASCII arrays.  60-column FASTA."""
    with open(path, "wb") as fh:
        for name, a in zip(names, seqs):
            fh.write((">%s\n" % name).encode())
            n = len(a)
            full = (n // width) * width
            body = np.empty((n // width, width + 1), dtype=np.uint8)
            body[:, :width] = a[:full].reshape(-1, width)
            body[:, width] = 10
            fh.write(body.tobytes())
            if full < n:
                fh.write(a[full:].tobytes() + b"\n")


def write_vcf(path, names, chrom, pos, ref, alt):
    with open(path, "w") as fh:
        fh.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\n")
        for c, p, r, a in zip(chrom, pos, ref, alt):
            fh.write("%s\t%d\t.\t%s\t%s\n" % (names[c], p, "ACGT"[r], "ACGT"[a]))


class ReferenceVEP(object):
    """The reference's AnalyzeSequences over a synthetic genome; ``score`` = one stock variant_effect_prediction call."""

    def __init__(self, names, seqs, workdir, sequence_length=1000, n_features=919, seed=0, threads=None):
        import torch
        self.threads = int(threads or os.cpu_count() or 1)
        torch.set_num_threads(self.threads)
        Genome, AnalyzeSequences, self._handler_mod, deepsea = _import_reference()
        os.makedirs(workdir, exist_ok=True)
        self.workdir, self.names = workdir, list(names)
        fasta = os.path.join(workdir, "genome.fa")
        write_fasta(names, seqs, fasta)
        torch.manual_seed(seed)
        model = deepsea.DeepSEA(sequence_length, n_features)
        weights = os.path.join(workdir, "deepsea_random_init.pth")
        torch.save(model.state_dict(), weights)
        self.az = AnalyzeSequences(model, weights, sequence_length=sequence_length,
                                   features=["f%d" % i for i in range(n_features)], batch_size=64, use_cuda=False,
                                   reference_sequence=Genome(fasta))
        self._calls = 0

    def score(self, chrom, pos, ref, alt, save_data=("abs_diffs", "logits"), write=False):
        """Returns dict name -> (n, F) float32 array when ``write`` is False, else the output directory."""
        import warnings
        self._calls += 1
        vcf = os.path.join(self.workdir, "step%d.vcf" % self._calls)
        write_vcf(vcf, self.names, chrom, pos, ref, alt)
        out_dir = os.path.join(self.workdir, "out%d" % self._calls)
        collected = {}
        stock = self._handler_mod.write_to_tsv_file
        if not write:
            def collect(data_across_features, info_cols, output_filepath):
                key = os.path.basename(output_filepath).rsplit(".", 1)[0].split("_", 1)[1]
                collected.setdefault(key, []).extend(data_across_features)
            self._handler_mod.write_to_tsv_file = collect
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                self.az.variant_effect_prediction(vcf, list(save_data), output_dir=out_dir)
        finally:
            self._handler_mod.write_to_tsv_file = stock
        if write:
            return out_dir
        return {k: np.concatenate(v) for k, v in collected.items()}
