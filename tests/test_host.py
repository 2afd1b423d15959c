"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/selene_b200.h
declares (no compute calls without a GPU), and the host logic around the kernels (FASTA image packing,
threshold parsing, VCF ingest, ref/alt window construction, BatchNorm folding, TSV formatting)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import selene_oracle as O  # noqa: E402
import sb_helpers as H  # noqa: E402


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "selene_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", txt)))


def test_cabi_exports_every_declared_symbol():
    from selene_b200 import _lib
    declared = _header_functions()
    assert len(declared) >= 25
    assert sorted(_lib.SIGNATURES.keys()) == declared          # the binding mirrors the header 1:1
    lib = _lib.load()                                           # raises if the .so is missing
    for name in declared:
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (sb_[a-z0-9_]+)\b", out))
    assert set(declared) <= exported
    assert lib.sb_version() >= 100


def test_no_cpu_fallback_in_product():
    """The product path never imports the oracle and refuses to run without CUDA."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "selene_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(dirpath, f)
    from selene_b200.predict.model_predict import AnalyzeSequences
    with pytest.raises(ValueError):
        AnalyzeSequences(None, sequence_length=1000, features=[], use_cuda=False)


def test_fasta_reader_and_image_packing(tmp_path):
    from selene_b200.sequences.genome import read_fasta, pack_image, codes_of_string
    g = H.synth_genome(n_chrom=3, length=1234)
    for width, fai in ((60, True), (50, True), (70, False)):
        path = str(tmp_path / ("g%d.fa" % width))
        H.write_fasta(g, path, width=width, with_fai=fai)
        recs = read_fasta(path)
        assert [n for n, _ in recs] == list(g.keys())
        for n, arr in recs:
            assert arr.tobytes().decode() == g[n]
    seqs = [np.frombuffer(s.encode(), np.uint8) for s in g.values()]
    packed, unk, upn, n_image, offs = pack_image(seqs)
    assert n_image % 64 == 0 and all(o % 64 == 0 for o in offs)
    for s, off in zip(g.values(), offs):
        idx = off + np.arange(len(s))
        code = (packed[idx >> 2] >> (2 * (idx & 3))) & 3
        u = (unk[idx >> 3] >> (idx & 7)) & 1
        n = (upn[idx >> 3] >> (idx & 7)) & 1
        exp = codes_of_string(s)
        assert np.array_equal(np.where(u == 1, 4, code), exp)
        assert np.array_equal(n == 1, np.frombuffer(s.encode(), np.uint8) == ord("N"))


def test_threshold_parsing_and_integer_thresholds():
    """genomic_features.py:144-193 and _genomic_features.pyx:39-40."""
    from selene_b200.targets.genomic_features import _define_feature_thresholds, integer_thresholds
    feats = ["a", "b", "c"]
    d, v = _define_feature_thresholds(0.5, feats)
    assert d == {"a": 0.5, "b": 0.5, "c": 0.5} and v.dtype == np.float32 and v.tolist() == [0.5] * 3
    d, v = _define_feature_thresholds({"default": 0.4, "b": 0.1}, feats)
    assert d["b"] == 0.1 and np.allclose(v, [0.4, 0.1, 0.4])
    d, v = _define_feature_thresholds(lambda f: 0.25 if f == "c" else 0.75, feats)
    assert np.allclose(v, [0.75, 0.75, 0.25])
    ks = np.arange(1001) / 1000.0
    assert np.array_equal(integer_thresholds(ks, 200), O.integer_thresholds(ks, 200))
    assert int(integer_thresholds([0.29], 200)[0]) == 57


class _FakeGenome(object):
    UNK_BASE = "N"

    def __init__(self, g):
        self.g = g

    def get_chrs(self):
        return sorted(self.g.keys())

    def coords_in_bounds(self, chrom, start, end):
        return O.check_coords({k: len(v) for k, v in self.g.items()}, chrom, start, end)

    def get_sequence_from_coords(self, chrom, start, end, strand='+', pad=False):
        return O.get_sequence_from_coords(self.g, chrom, start, end, strand, pad)


def test_read_vcf_file(tmp_path):
    """_variant_effect_prediction.py:13-134: chrom normalisation, '-' ref, multi-alt split, strand
    column, out-of-bounds rows -> .NA file."""
    from selene_b200.predict._variant_effect_prediction import read_vcf_file
    g = {"chr1": "ACGT" * 600, "chrM": "ACGT" * 600}
    vcf = tmp_path / "v.vcf"
    vcf.write_text("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tSTRAND\n"
                   "1\t1200\trs1\tA\tC,G\t+\n"
                   "CHR1\t1300\trs2\t-\tT\t-\n"
                   "chrMT\t1250\trs3\tG\tA.T\t.\n"
                   "chr1\t10\trs4\tA\tC\t+\n"
                   "chr9\t1200\trs5\tA\tC\t+\n"
                   "short\tline\n")
    na = str(tmp_path / "na.txt")
    v = read_vcf_file(str(vcf), strand_index=5, output_NAs_to_file=na, seq_context=(500, 500),
                      reference_sequence=_FakeGenome(g))
    assert v == [("chr1", 1200, "rs1", "A", "C", "+"), ("chr1", 1200, "rs1", "A", "G", "+"),
                 ("chr1", 1300, "rs2", "", "T", "-"), ("chrM", 1250, "rs3", "G", "A", "+"),
                 ("chrM", 1250, "rs3", "G", "T", "+")]
    assert len(open(na).read().strip().split("\n")) == 3
    with pytest.raises(ValueError):
        bad = tmp_path / "bad.vcf"
        bad.write_text("#CHROM\tPOSITION\tID\tREF\tALT\n1\t1\t.\tA\tC\n")
        read_vcf_file(str(bad), reference_sequence=_FakeGenome(g))


def test_ref_alt_strings_match_oracle_for_snvs():
    from selene_b200.predict._variant_effect_prediction import build_ref_alt_strings, _get_ref_idxs
    assert _get_ref_idxs(1000, 1) == (499, 500) and _get_ref_idxs(1001, 1) == (500, 501)
    assert _get_ref_idxs(1000, 3) == (498, 501) == O.get_ref_idxs(1000, 3)
    g = H.synth_genome(n_chrom=2, length=5000)
    G = _FakeGenome(g)
    rng = np.random.default_rng(4)
    L = 200
    sr, er = O.radii(L)
    for _ in range(60):
        c = list(g.keys())[rng.integers(0, 2)]
        pos = int(rng.integers(sr + 1, len(g[c]) - er))
        ref = "ACGTN"[rng.integers(0, 5)]
        alt = "ACGTN"[rng.integers(0, 5)]
        strand = "+-"[rng.integers(0, 2)]
        rs, als, m, u = build_ref_alt_strings((c, pos, ".", ref, alt, strand), G, L, sr, er)
        r, a, m2, u2 = O.snv_ref_alt_encodings(g, c, pos, ref, alt, strand, L)
        assert np.array_equal(O.sequence_to_encoding(rs), r) and np.array_equal(O.sequence_to_encoding(als), a)
        assert m == m2 and u == u2
    # insertion / deletion keep the model input length
    for ref, alt in (("A", "ACGT"), ("ACGT", "A"), ("ACG", "TTT"), ("A", "-")):
        rs, als, _, _ = build_ref_alt_strings(("chr1", 2500, ".", ref, alt, "+"), G, L, sr, er)
        assert len(rs) == L and len(als) == L


def test_is_positive_row_reference_kats():
    """targets/tests/test_genomic_features.py:63-85 (the three known answers of _is_positive_row)."""
    from selene_b200.targets.genomic_features import _is_positive_row
    assert not _is_positive_row(16150, 16351, 16110, 16190, 0.50)
    assert _is_positive_row(16110, 16309, 16110, 16190, 0.40)
    assert _is_positive_row(16110, 16311, 16110, 16290, 0.80)


def test_batchnorm_folding_matches_torch_eval():
    """nets.layers_from_module: eval-mode BatchNorm folded into the following layer."""
    import torch
    import torch.nn as nn
    from selene_b200.nets import layers_from_module
    torch.manual_seed(1)
    m = nn.Sequential(nn.Conv1d(4, 6, 3), nn.ReLU(), nn.Conv1d(6, 8, 3), nn.ReLU(), nn.MaxPool1d(2, 2),
                      nn.BatchNorm1d(8), nn.Conv1d(8, 5, 2), nn.ReLU(), nn.Dropout(0.3))
    cls = nn.Sequential(nn.Linear(5 * 7, 9), nn.ReLU(), nn.BatchNorm1d(9), nn.Linear(9, 3), nn.Sigmoid())
    for bn in (m[5], cls[2]):
        bn.running_mean.uniform_(-0.5, 0.5)
        bn.running_var.uniform_(0.5, 1.5)
        bn.weight.data.uniform_(0.5, 1.5)
        bn.bias.data.uniform_(-0.5, 0.5)

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv_net, self.classifier = m, cls

        def forward(self, x):
            o = self.conv_net(x)
            return self.classifier(o.view(o.size(0), -1))
    net = Net().eval()
    x = np.random.default_rng(0).random((5, 20, 4)).astype(np.float32)
    with torch.no_grad():
        exp = net(torch.tensor(x).transpose(1, 2)).numpy()
    got = O.stack_forward(x, layers_from_module(net))
    assert np.abs(got - exp).max() < 1e-5


def test_indel_windows_match_reference_goldens():
    """Insertions, deletions and MNVs: build_ref_alt_strings vs ref/alt tensors produced by the
    unmodified reference (_process_alt / _handle_standard_ref, tests/golden/make_golden.py)."""
    from selene_b200.predict._variant_effect_prediction import build_ref_alt_strings
    gold = np.load(os.path.join(HERE, "golden", "reference_outputs.npz"))
    g = H.synth_genome()
    G = _FakeGenome(g)
    L = 200
    for i, line in enumerate(gold["indel_variants"]):
        c, pos, ref, alt, strand = str(line).split("\t")
        rs, als, m, u = build_ref_alt_strings((c, int(pos), ".", ref, alt, strand), G, L, L // 2, L // 2)
        assert np.array_equal(O.sequence_to_encoding(rs), gold["indel_ref_onehot"][i]), line
        assert np.array_equal(O.sequence_to_encoding(als), gold["indel_alt_onehot"][i]), line
        assert m == bool(gold["indel_match"][i]) and u == bool(gold["indel_unk"][i])


def test_count_unknown_host_prefix():
    """Genome.count_unknown (the sampler's rejection test) against a direct count, without a GPU."""
    from selene_b200.sequences.genome import Genome
    g = H.synth_genome(n_chrom=1, length=3000, n_frac=0.1)
    G = Genome.__new__(Genome)
    seq = g["chr1"]
    G._initialized = True
    G._seqs = {"chr1": np.frombuffer(seq.encode(), np.uint8)}
    G.len_chrs = {"chr1": len(seq)}
    rng = np.random.default_rng(1)
    for _ in range(200):
        s = int(rng.integers(-50, len(seq)))
        e = s + int(rng.integers(1, 400))
        inner = seq[max(s, 0):min(e, len(seq))]
        exp = sum(ch not in "ACGTacgt" for ch in inner) + max(0, -s) + max(0, e - len(seq))
        assert G.count_unknown("chr1", s, e) == exp


def test_hdf5_writer_or_clear_import_error(tmp_path):
    """handler.py:205-218: dataset ``data`` float64 (n, F).  With h5py the file is written and read back; without it
    (this image) ``output_format='hdf5'`` must fail with an ImportError that names the package, not crash later."""
    from selene_b200.predict._writers import write_hdf5
    data = np.arange(12, dtype=np.float32).reshape(3, 4) / 7
    path = str(tmp_path / "scores.h5")
    try:
        import h5py
    except ImportError:
        with pytest.raises(ImportError, match="h5py"):
            write_hdf5(path, data)
        return
    write_hdf5(path, data)
    with h5py.File(path, "r") as fh:
        got = fh["data"][()]
    assert got.dtype == np.float64 and np.array_equal(got, data.astype(np.float64))
