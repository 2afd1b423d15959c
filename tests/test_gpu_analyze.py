"""GPU end-to-end tests of the reference-facing entry points: AnalyzeSequences.in_silico_mutagenesis
and .variant_effect_prediction write the reference's files with the reference's numbers (oracle)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, HERE)
from oracle import selene_oracle as O  # noqa: E402
import sb_helpers as H  # noqa: E402

pytestmark = pytest.mark.gpu
FEATS = ["f%d" % i for i in range(919)]


def _deepsea():
    import torch
    import torch.nn as nn

    class DeepSEA(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv_net = nn.Sequential(
                nn.Conv1d(4, 320, 8), nn.ReLU(inplace=True), nn.MaxPool1d(4, 4), nn.Dropout(p=0.2),
                nn.Conv1d(320, 480, 8), nn.ReLU(inplace=True), nn.MaxPool1d(4, 4), nn.Dropout(p=0.2),
                nn.Conv1d(480, 960, 8), nn.ReLU(inplace=True), nn.Dropout(p=0.5))
            self.classifier = nn.Sequential(nn.Linear(960 * 53, 919), nn.ReLU(inplace=True),
                                            nn.Linear(919, 919), nn.Sigmoid())
    torch.manual_seed(0)
    return DeepSEA()


def _read_tsv(path, n_id):
    lines = open(path).read().rstrip("\n").split("\n")
    header = lines[0].split("\t")
    ids = [tuple(l.split("\t")[:n_id]) for l in lines[1:]]
    data = np.array([[float(x) for x in l.split("\t")[n_id:]] for l in lines[1:]])
    return header, ids, data


def test_ism_files_match_oracle(tmp_path):
    from selene_b200.predict.model_predict import AnalyzeSequences
    from selene_b200.sequences import Genome
    az = AnalyzeSequences(_deepsea(), sequence_length=1000, features=FEATS, reference_sequence=Genome,
                          precision="fp32")
    r = np.random.default_rng(0)
    seq = "".join(np.array(list("ACGT"))[r.integers(0, 4, 990)].tolist())
    seq = seq[:100] + "N" + seq[101:200] + "acgt" + seq[204:]       # one unknown, some lowercase, needs padding
    prefix = str(tmp_path / "out" / "ism")
    az.in_silico_mutagenesis(seq, ["abs_diffs", "logits", "predictions"], output_path_prefix=prefix)
    padded = ("N" * 5 + seq + "N" * 5).upper()
    muts, absd, logits, base = O.ism_scores(padded, O.deepsea_layers(1000, 919, seed=0))
    exp_ids = [(str(m[0][0]), padded[m[0][0]], m[0][1]) for m in muts]
    for name, exp in (("abs_diffs", absd), ("logits", logits)):
        header, ids, data = _read_tsv(prefix + "_%s.tsv" % name, 3)
        assert header == ["pos", "ref", "alt"] + FEATS and ids == exp_ids
        assert data.shape == exp.shape
        assert np.abs(data - exp).max() <= 1e-5 + 6e-3 * np.abs(exp).max()    # "{:.2e}" keeps 3 digits
    header, ids, data = _read_tsv(prefix + "_predictions.tsv", 3)
    assert ids[0] == ("input_sequence", "NA", "NA") and ids[1:] == exp_ids
    assert np.abs(data[0] - base[0]).max() <= 6e-3


def test_vep_files_match_oracle(tmp_path):
    from selene_b200.predict.model_predict import AnalyzeSequences
    from selene_b200.sequences import Genome
    g = H.synth_genome()
    fa = str(tmp_path / "g.fa")
    H.write_fasta(g, fa)
    Genome.update_bases_order(['A', 'C', 'G', 'T'])
    G = Genome(fa)
    az = AnalyzeSequences(_deepsea(), sequence_length=1000, features=FEATS, reference_sequence=G, precision="fp32")
    rng = np.random.default_rng(8)
    rows, variants = [], []
    for i in range(20):
        c = sorted(g.keys())[rng.integers(0, 3)]
        pos = int(rng.integers(600, len(g[c]) - 600))
        ref = g[c][pos - 1].upper()
        ref = ref if ref in "ACGT" and rng.random() > 0.2 else "ACGT"[rng.integers(0, 4)]
        alt = "ACGT"[rng.integers(0, 4)]
        strand = "+-"[rng.integers(0, 2)]
        rows.append("%s\t%d\trs%d\t%s\t%s\t%s\n" % (c, pos, i, ref, alt, strand))
        variants.append((c, pos, ref, alt, strand))
    rows.append("chr1\t700\tmnv\tAC\tGT\t+\n")              # MNV: host-built windows
    variants.append(("chr1", 700, "AC", "GT", "+"))
    rows.append("chr1\t10\toob\tA\tC\t+\n")                  # out of bounds -> .NA file
    vcf = str(tmp_path / "in.vcf")
    open(vcf, "w").write("#CHROM\tPOS\tID\tREF\tALT\tSTRAND\n" + "".join(rows))
    out_dir = str(tmp_path / "vep")
    with pytest.warns(UserWarning):
        az.variant_effect_prediction(vcf, ["abs_diffs", "logits", "predictions"], output_dir=out_dir, strand_index=5)
    absd, logits, match, unk = O.vep_scores(g, variants, O.deepsea_layers(1000, 919, seed=0), 1000)
    header, ids, data = _read_tsv(os.path.join(out_dir, "in_abs_diffs.tsv"), 8)
    assert header[:8] == ["chrom", "pos", "name", "ref", "alt", "strand", "ref_match", "contains_unk"]
    assert [i[6] for i in ids] == [str(bool(m)) for m in match]
    assert [i[7] for i in ids] == [str(bool(u)) for u in unk]
    assert [(i[0], int(i[1]), i[3], i[4], i[5]) for i in ids] == variants
    assert np.abs(data - absd).max() <= 1e-5 + 6e-3 * np.abs(absd).max()
    _, _, data = _read_tsv(os.path.join(out_dir, "in_logits.tsv"), 8)
    assert np.abs(data - logits).max() <= 1e-4 + 6e-3 * np.abs(logits).max()
    assert os.path.exists(os.path.join(out_dir, "in.ref_predictions.tsv"))
    assert os.path.exists(os.path.join(out_dir, "in.alt_predictions.tsv"))
    assert open(os.path.join(out_dir, "in.NA")).read().count("\n") == 1


def test_ism_with_non_strand_specific_model():
    """AnalyzeSequences on a model wrapped in NonStrandSpecific (every manuscript config): ISM scores
    equal those of the torch wrapper evaluated on the mutated one-hot sequences."""
    import torch
    import sb_models as M
    from selene_b200.predict.model_predict import AnalyzeSequences
    from selene_b200.sequences import Genome
    Genome.update_bases_order(['A', 'C', 'G', 'T'])
    torch.manual_seed(4)
    wrapped = M.NonStrandSpecific(M.DeepSEA(1000, 919), mode="mean").eval()
    az = AnalyzeSequences(wrapped, sequence_length=1000, features=FEATS, reference_sequence=Genome, precision="fp32")
    r = np.random.default_rng(2)
    seq = "".join(np.array(list("ACGT"))[r.integers(0, 4, 1000)].tolist())
    (pos, alt), out, base = az.ism_scores(seq, ("abs_diffs",), start_position=480, end_position=500)
    assert len(pos) == 60
    enc = O.sequence_to_encoding(seq)
    rows = np.repeat(enc[None], 60, axis=0)
    for i in range(60):
        rows[i, pos[i]] = 0
        rows[i, pos[i], alt[i]] = 1
    with torch.no_grad():
        exp_base = wrapped(torch.tensor(enc[None]).transpose(1, 2)).numpy()
        exp = wrapped(torch.tensor(rows).transpose(1, 2)).numpy()
    assert np.abs(base - exp_base).max() <= 1e-5
    assert np.abs(out["abs_diffs"] - np.abs(exp_base - exp)).max() <= 2e-5


def test_get_predictions_fasta_bed_and_ism_from_file(tmp_path):
    """model_predict.py:347-597,801-950: predictions for a raw sequence / FASTA / BED input and ISM over a FASTA
    file write the reference's files; numbers against torch on the oracle's encodings."""
    import torch
    from selene_b200.predict.model_predict import AnalyzeSequences
    from selene_b200.sequences import Genome
    g = H.synth_genome(n_chrom=2, length=6000)
    fa = str(tmp_path / "g.fa")
    H.write_fasta(g, fa)
    Genome.update_bases_order(['A', 'C', 'G', 'T'])
    G = Genome(fa)
    import sb_models as M
    torch.manual_seed(0)
    model = M.DeepSEA(1000, 919).eval()
    az = AnalyzeSequences(model, sequence_length=1000, features=FEATS, reference_sequence=G, precision="fp32", batch_size=4)

    def torch_pred(seqs):
        x = torch.tensor(np.stack([O.sequence_to_encoding(s) for s in seqs])).transpose(1, 2)
        with torch.no_grad():
            return model(x).numpy()

    # ---- raw sequence (short: padded with N on both sides)
    r = np.random.default_rng(3)
    raw = "".join(np.array(list("ACGTacgtN"))[r.integers(0, 9, 700)].tolist())
    got = az.get_predictions(raw)
    assert got.shape == (1, 919) and np.abs(got - torch_pred([az._pad_or_truncate_sequence(raw)])).max() <= 1e-5

    # ---- FASTA: records shorter, equal and longer than the model input
    recs = {"s_short": raw, "s_exact": "".join(np.array(list("ACGT"))[r.integers(0, 4, 1000)].tolist()),
            "s_long": "".join(np.array(list("ACGT"))[r.integers(0, 4, 1300)].tolist())}
    qfa = str(tmp_path / "q.fa")
    with open(qfa, "w") as fh:
        for k, v in recs.items():
            fh.write(">%s\n" % k)
            for i in range(0, len(v), 60):
                fh.write(v[i:i + 60] + "\n")
    out = str(tmp_path / "pf")
    az.get_predictions(qfa, out)
    header, ids, data = _read_tsv(os.path.join(out, "q_predictions.tsv"), 2)
    assert header == ["index", "name"] + FEATS and ids == [(str(i), k) for i, k in enumerate(recs)]
    exp = torch_pred([az._pad_or_truncate_sequence(v) for v in recs.values()])
    assert np.abs(data - exp).max() <= 6e-3 * max(1.0, np.abs(exp).max())      # "{:.2e}" keeps 3 digits

    # ---- BED: centred windows, strand column, one row out of bounds, one unknown contig
    c0, c1 = sorted(g.keys())[:2]
    bed = str(tmp_path / "r.bed")
    rows = [(c0, 1500, 1700, "+"), (c1, 3000, 3001, "-"), (c0, 10, 20, "+"), ("chrNope", 5, 9, "+"), (c1, 2500, 2600, ".")]
    open(bed, "w").write("".join("%s\t%d\t%d\t%s\n" % r_ for r_ in rows))
    out = str(tmp_path / "pb")
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        az.get_predictions_for_bed_file(bed, out, strand_index=3)
    header, ids, data = _read_tsv(os.path.join(out, "r_predictions.tsv"), 6)
    assert header[:6] == ["index", "chrom", "start", "end", "strand", "contains_unk"]
    assert [i[0] for i in ids] == ["0", "1", "4"]
    assert len(open(os.path.join(out, "r.NA")).read().strip().split("\n")) == 2
    seqs = []
    for (c, s, e, st) in (rows[0], rows[1], rows[4]):
        mid = s + (e - s) // 2
        seqs.append(O.get_sequence_from_coords(g, c, mid - 500, mid + 500, st if st != "." else "+", True))
    exp = torch_pred(seqs)
    assert np.abs(data - exp).max() <= 6e-3 * max(1.0, np.abs(exp).max())
    assert [i[5] for i in ids] == [str("N" in s) for s in seqs]

    # ---- ISM over a FASTA file = in_silico_mutagenesis per record
    out = str(tmp_path / "ismf")
    az.in_silico_mutagenesis_from_file(qfa, ["abs_diffs"], out, start_position=495, end_position=505)
    az.in_silico_mutagenesis(recs["s_long"], ["abs_diffs"], output_path_prefix=str(tmp_path / "single" / "x"),
                             start_position=495, end_position=505)
    a = open(os.path.join(out, "s_long_abs_diffs.tsv")).read()
    b = open(str(tmp_path / "single" / "x_abs_diffs.tsv")).read()
    assert a == b and len(a.split("\n")) == 32
    assert sorted(os.listdir(out)) == ["s_exact_abs_diffs.tsv", "s_long_abs_diffs.tsv", "s_short_abs_diffs.tsv"]
    az.in_silico_mutagenesis_from_file(qfa, ["abs_diffs"], str(tmp_path / "ismi"), use_sequence_name=False,
                                       start_position=0, end_position=2)
    assert sorted(os.listdir(str(tmp_path / "ismi"))) == ["0_abs_diffs.tsv", "1_abs_diffs.tsv", "2_abs_diffs.tsv"]


def test_ism_from_file_honours_mutate_n_bases(tmp_path):
    """model_predict.py:907-912: ``in_silico_mutagenesis_from_file`` passes ``mutate_n_bases`` on; ids are
    ``pos;pos`` / ``ref;ref`` / ``alt;alt`` (_in_silico_mutagenesis.py:146-170), order = combinations x products."""
    from selene_b200.predict.model_predict import AnalyzeSequences
    from selene_b200.sequences import Genome
    Genome.update_bases_order(['A', 'C', 'G', 'T'])
    az = AnalyzeSequences(_deepsea(), sequence_length=1000, features=FEATS, reference_sequence=Genome,
                          precision="fp32", batch_size=64)
    r = np.random.default_rng(12)
    seq = "".join(np.array(list("ACGT"))[r.integers(0, 4, 1000)].tolist())
    fa = str(tmp_path / "two.fa")
    open(fa, "w").write(">rec1 first token is the name\n%s\n" % seq)
    out = str(tmp_path / "o")
    az.in_silico_mutagenesis_from_file(fa, ["abs_diffs", "logits", "predictions"], out, mutate_n_bases=2,
                                       start_position=497, end_position=503)
    layers = O.deepsea_layers(1000, 919, seed=0)
    muts = O.in_silico_mutagenesis_sequences(seq, 2, start_position=497, end_position=503)
    assert len(muts) == 15 * 9
    enc = O.sequence_to_encoding(seq)
    base = O.stack_forward(enc[None], layers)
    absd, logits = [], []
    for i in range(0, len(muts), 64):
        batch = np.stack([O.mutate_sequence(enc, m) for m in muts[i:i + 64]])
        pred = O.stack_forward(batch, layers)
        absd.append(O.abs_diff_score(pred, base))
        logits.append(O.logit_score(pred, base))
    absd, logits = np.concatenate(absd), np.concatenate(logits)
    exp_ids = [(";".join(str(p) for p, _ in m), ";".join(seq[p] for p, _ in m), ";".join(b for _, b in m)) for m in muts]
    for name, exp in (("abs_diffs", absd), ("logits", logits)):
        header, ids, data = _read_tsv(os.path.join(out, "rec1_%s.tsv" % name), 3)
        assert ids == exp_ids
        assert np.abs(data - exp).max() <= 1e-5 + 6e-3 * np.abs(exp).max()
    header, ids, data = _read_tsv(os.path.join(out, "rec1_predictions.tsv"), 3)
    assert ids[0] == ("input_sequence", "NA", "NA") and ids[1:] == exp_ids


def test_vep_two_processes_write_identical_files(tmp_path):
    """variants sharded over ranks, every rank writes its rows into the shared files at its byte offset
    ("shards"), or rank 0 gathers the tiles over NCCL and writes ("gather").  Both must be byte-identical to the
    single-process files.  Needs >= 2 GPUs (run with gpurun --gpus 2); skipped otherwise."""
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = os.path.join(HERE, "multi_gpu_vep_check.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    single = subprocess.run([sys.executable, script, str(tmp_path), "single"], env=env, capture_output=True, text=True)
    assert single.returncode == 0, single.stdout[-2000:] + single.stderr[-2000:]
    for mode in ("shards", "gather"):
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                            "--master-addr", "127.0.0.1", "--master-port", "29533", script, str(tmp_path), mode],
                           env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        for f in sorted(os.listdir(os.path.join(str(tmp_path), "single"))):
            a = open(os.path.join(str(tmp_path), "single", f), "rb").read()
            b = open(os.path.join(str(tmp_path), mode, f), "rb").read()
            assert a == b, "%s differs between the single-process and the 2-rank %s run" % (f, mode)
