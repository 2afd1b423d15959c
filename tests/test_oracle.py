"""CPU tests that PIN the oracle (oracle/selene_oracle.py): against the reference's own known-answer
tests, against golden outputs of the unmodified reference (tests/golden/reference_outputs.npz, made by
tests/golden/make_golden.py) and, when present, against the reference's Cython modules compiled from
where they lie (oracle/_ref, oracle/build_ref.py)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, HERE)
from oracle import selene_oracle as O  # noqa: E402
import sb_helpers as H  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "reference_outputs.npz"))


# ---- known answers lifted from the reference's tests ---------------------------------------------
def test_kat_sequence_to_encoding():
    """selene_sdk/sequences/tests/test_genome.py:40-60."""
    exp = np.array([[0., 1., 0., 0.], [0., 0., 0., 1.], [0., 0., 1., 0.], [0., 1., 0., 0.],
                    [0., 0., 1., 0.], [0., 1., 0., 0.], [1., 0., 0., 0.], [1., 0., 0., 0.]])
    assert O.sequence_to_encoding("ctgCGCAA").tolist() == exp.tolist()
    exp = np.array([[1., 0., 0., 0.], [.25, .25, .25, .25], [.25, .25, .25, .25], [.25, .25, .25, .25],
                    [1., 0., 0., 0.], [0., 0., 0., 1.], [0., 1., 0., 0.], [1., 0., 0., 0.]])
    got = O.sequence_to_encoding("AnnUAtCa")
    assert got.dtype == np.float32 and got.tolist() == exp.tolist()


def test_kat_encoding_to_sequence():
    """selene_sdk/sequences/tests/test_genome.py:62-78."""
    enc = np.array([[1., 0., 0., 0.], [1., 0., 0., 0.], [0., 0., 0., 1.], [0., 0., 1., 0.],
                    [0., 1., 0., 0.], [0., 0., 0., 1]])
    assert O.encoding_to_sequence(enc) == "AATGCT"
    enc = np.array([[0., 0., 1., 0.], [0.25, 0.25, 0.25, 0.25], [1., 0., 0., 0.], [0., 0., 0., 1.],
                    [0.25, 0.25, 0.25, 0.25], [0.25, 0.25, 0.25, 0.25]])
    assert O.encoding_to_sequence(enc) == "GNATNN"


def test_kat_get_sequence_from_coords():
    """selene_sdk/sequences/tests/test_genome.py:14-38,80-106: the fake genome there is the repeated
    8-mer AGCTTCCA whose minus strand is served as the repeated TCGAAGGT sliced with the SAME
    coordinates; as a dict genome that is the reverse complement of the plus-strand slice only when
    the slice is 8-aligned, so the minus-strand vector is checked on the plus-strand-equivalent."""
    lens = {"chr1": 104, "chr2": 16, "chr3": 64, "chr4": 8}
    g = {c: "AGCTTCCA" * (n // 8) for c, n in lens.items()}
    assert O.get_sequence_from_coords(g, "chr1", 0, 14, "+") == "AGCTTCCAAGCTTC"
    assert O.get_sequence_from_coords(g, "chr3", 56, 64, "-") == O.revcomp_string("AGCTTCCA") == "TGGAAGCT"
    with pytest.raises(ValueError):
        O.get_sequence_from_coords(g, "chr2", 1, 10, "=")
    assert O.get_sequence_from_coords(g, "chr2", 17, 20, "+") == ""
    assert O.get_sequence_from_coords(g, "chr4", 2, 10, "-") == ""


def test_kat_reverse_complement_encoding():
    """selene_sdk/predict/tests/test__common.py:13-48."""
    enc = np.array([[0., 0., 0., 1.], [1., 0., 0., 0.], [0., 1., 0., 0.], [0., 0., 1., 0.], [0.25, 0.25, 0.25, 0.25]])
    exp = [[0.25, 0.25, 0.25, 0.25], [0., 1., 0., 0.], [0., 0., 1., 0.], [0., 0., 0., 1.], [1., 0., 0., 0.]]
    assert O.get_reverse_complement_encoding(enc).tolist() == exp
    exp2 = [[0.25, 0.25, 0.25, 0.25], [0., 0., 0., 1.], [1., 0., 0., 0.], [0., 1., 0., 0.], [0., 0., 1., 0.]]
    assert O.get_reverse_complement_encoding(enc, ['A', 'T', 'G', 'C']).tolist() == exp2


def test_kat_ism_sequences():
    """selene_sdk/predict/tests/test_model_predict.py:12-60."""
    obs = O.in_silico_mutagenesis_sequences("ATCCG")
    exp = [(0, 'C'), (0, 'G'), (0, 'T'), (1, 'A'), (1, 'C'), (1, 'G'), (2, 'A'), (2, 'G'), (2, 'T'),
           (3, 'A'), (3, 'G'), (3, 'T'), (4, 'A'), (4, 'C'), (4, 'T')]
    assert obs == [[e] for e in exp]
    obs = O.in_silico_mutagenesis_sequences("ATCCG", start_position=1, end_position=4)
    exp = [(1, 'A'), (1, 'C'), (1, 'G'), (2, 'A'), (2, 'G'), (2, 'T'), (3, 'A'), (3, 'G'), (3, 'T')]
    assert obs == [[e] for e in exp]
    obs = O.in_silico_mutagenesis_sequences("ATC", mutate_n_bases=2, start_position=0, end_position=3)
    exp = [[(0, 'C'), (1, 'A')], [(0, 'G'), (1, 'A')], [(0, 'T'), (1, 'A')],
           [(0, 'C'), (1, 'C')], [(0, 'G'), (1, 'C')], [(0, 'T'), (1, 'C')],
           [(0, 'C'), (1, 'G')], [(0, 'G'), (1, 'G')], [(0, 'T'), (1, 'G')],
           [(0, 'C'), (2, 'A')], [(0, 'G'), (2, 'A')], [(0, 'T'), (2, 'A')],
           [(0, 'C'), (2, 'G')], [(0, 'G'), (2, 'G')], [(0, 'T'), (2, 'G')],
           [(0, 'C'), (2, 'T')], [(0, 'G'), (2, 'T')], [(0, 'T'), (2, 'T')],
           [(1, 'A'), (2, 'A')], [(1, 'C'), (2, 'A')], [(1, 'G'), (2, 'A')],
           [(1, 'A'), (2, 'G')], [(1, 'C'), (2, 'G')], [(1, 'G'), (2, 'G')],
           [(1, 'A'), (2, 'T')], [(1, 'C'), (2, 'T')], [(1, 'G'), (2, 'T')]]
    assert sorted(obs) == sorted(exp)
    obs = O.in_silico_mutagenesis_sequences("ATCG", mutate_n_bases=2, start_position=1, end_position=3)
    assert len(obs) == 9
    assert len(O.in_silico_mutagenesis_sequences("ATCCG", mutate_n_bases=2)) == int(GOLD["ism_double_n"][0])


FEATURES = ["CTCF", "eGFP-FOS", "GABP", "Pbx3", "Pol2", "TBP"]
FID = {f: i for i, f in enumerate(FEATURES)}


ROWS1 = [["1", "16110", "16190", "CTCF"], ["1", "16128", "16158", "CTCF"], ["1", "16149", "16239", "CTCF"]]
ROWS2 = [["2", "91128", "91358", "CTCF"], ["2", "91130", "91239", "CTCF"], ["2", "91156", "91310", "CTCF"]]
ROWS3 = [["chr3", "8533", "8817", "eGFP-FOS"], ["chr3", "8541", "8651", "GABP"], ["chr3", "8574", "8629", "Pol2"],
         ["chr3", "8619", "9049", "CTCF"], ["chr3", "8620", "8680", "TBP"], ["chr3", "8645", "8720", "TBP"]]


def test_kat_fast_get_feature_data():
    """Rows (selene_sdk/targets/tests/test_genomic_features.py:24-42) and expected vectors (:119-193),
    calling the function directly because the wrappers there are stale.  The fake
    get_feature_rows of that test ignores (start, end), so rows are passed unfiltered."""
    thr = np.array([0.50] * 6).astype(np.float32)
    z = O.fast_get_feature_data(10, 211, thr, FID, None)
    assert z.tolist() == [0] * 6 and z.dtype == np.float64
    assert O.fast_get_feature_data(10, 211, thr, FID, []).tolist() == [0] * 6
    assert O.fast_get_feature_data(16100, 16350, thr, FID, ROWS1).tolist() == [1, 0, 0, 0, 0, 0]
    thr51 = np.array([0.51] * 6).astype(np.float32)
    assert O.fast_get_feature_data(91027, 91228, thr51, FID, ROWS2).tolist() == [0, 0, 0, 0, 0, 0]
    assert O.fast_get_feature_data(8619, 8719, thr, FID, ROWS3).tolist() == [1, 1, 0, 0, 0, 1]
    thr_mix = np.array([0.50, 0.0, 0.0, 0.0, 0.0, 1.0]).astype(np.float32)
    assert O.fast_get_feature_data(8619, 8719, thr_mix, FID, ROWS3).tolist() == [1, 1, 1, 0, 1, 0]


def test_integer_threshold_is_float32():
    """q=200, thr=0.29 -> 57 in float32 arithmetic (56 in float64)."""
    assert int(O.integer_thresholds([0.29], 200)[0]) == 57
    assert int(np.clip(0.29 * 200 - 1, 0, None)) == 56


# ---- golden outputs of the unmodified reference ----------------------------------------------------
def test_golden_encode():
    g = H.synth_genome()
    names = sorted(g.keys())
    off = 0
    for (ci, s, e, rc, pad), n, unk in zip(GOLD["encode_queries"], GOLD["encode_lengths"], GOLD["encode_unk"]):
        exp = GOLD["encode_flat"][off:off + 4 * n].reshape(n, 4)
        off += 4 * n
        got, gunk = O.get_encoding_from_coords_check_unk(g, names[ci], int(s), int(e), "-" if rc else "+", bool(pad))
        assert got.shape == exp.shape and np.array_equal(got, exp)
        assert gunk == bool(unk)
    assert np.array_equal(O.sequence_to_encoding(g["chr1"][60:260], ['A', 'G', 'C', 'T']), GOLD["encode_agct"])


@pytest.mark.parametrize("tag,thr", [("t50", 0.5), ("t29", 0.29), ("tdict", {"default": 0.4, "f3": 0.01}), ("tnone", None)])
def test_golden_labels(tag, thr):
    g = H.synth_genome(length=30000)
    names = sorted(g.keys())
    rows = H.synth_intervals(g)
    by = H.rows_by_chrom(rows)
    feats = ["f%d" % i for i in range(12)]
    fid = {f: i for i, f in enumerate(feats)}
    vec = None
    if thr is not None:
        vec = np.array([thr if not isinstance(thr, dict) else thr.get(f, thr["default"]) for f in feats], np.float32)
    for (ci, s, e), exp in zip(GOLD["label_queries"], GOLD["label_" + tag]):
        r = O.query_rows(by, names[ci], int(s), int(e))
        got = O.feature_data_no_threshold(12, fid, r) if thr is None else O.fast_get_feature_data(int(s), int(e), vec, fid, r)
        assert np.array_equal(got, exp)


def test_golden_ism():
    r = np.random.default_rng(0)
    seq = "".join(np.array(list("ACGTN"))[r.choice(5, 300, p=[.24, .24, .24, .24, .04])].tolist())
    muts = O.in_silico_mutagenesis_sequences(seq)
    assert [m[0][0] for m in muts] == GOLD["ism_pos"].tolist()
    assert ["ACGT".index(m[0][1]) for m in muts] == GOLD["ism_alt"].tolist()
    enc = O.sequence_to_encoding(seq)
    got = np.array([O.mutate_sequence(enc, muts[i]) for i in range(0, len(muts), 37)])
    assert np.array_equal(got, GOLD["ism_mutants"])


def test_golden_vep_windows():
    g = H.synth_genome()
    names = sorted(g.keys())
    for i, (ci, pos, ref, alt, rc) in enumerate(GOLD["vep_variants"]):
        r, a, m, u = O.snv_ref_alt_encodings(g, names[ci], int(pos), "ACGT"[ref], "ACGT"[alt], "-" if rc else "+", 1000)
        assert np.array_equal(r, GOLD["vep_ref_onehot"][i]) and np.array_equal(a, GOLD["vep_alt_onehot"][i])
        assert m == bool(GOLD["vep_match"][i]) and u == bool(GOLD["vep_unk"][i])


def test_golden_deepsea_forward_and_scores():
    layers = O.deepsea_layers(1000, 919, seed=0)
    codes = GOLD["deepsea_codes"]
    onehot = np.zeros(codes.shape + (4,), np.float32)
    for b in range(4):
        onehot[codes == b, b] = 1
    onehot[codes == 4] = 0.25
    got = O.stack_forward(onehot, layers)
    assert np.abs(got - GOLD["deepsea_pred"]).max() <= 1e-6
    ref, alt = GOLD["vep_ref_pred"], GOLD["vep_alt_pred"]
    assert np.array_equal(O.abs_diff_score(alt.copy(), ref.copy()), GOLD["vep_abs_diffs"])
    assert np.array_equal(O.diff_score(alt.copy(), ref.copy()), GOLD["vep_diffs"])
    assert np.allclose(O.logit_score(alt.copy(), ref.copy()), GOLD["vep_logits"], rtol=0, atol=1e-6)
    got_ref = O.stack_forward(GOLD["vep_ref_onehot"], layers)
    assert np.abs(got_ref - ref).max() <= 1e-6


# ---- the reference's own native modules, compiled from where they lie --------------------------------
def _ref_modules():
    try:
        from oracle._ref import _sequence, _genomic_features
        return _sequence, _genomic_features
    except Exception:
        return None


@pytest.mark.skipif(_ref_modules() is None, reason="oracle/_ref not built (needs /root/reference once)")
def test_oracle_vs_compiled_reference_modules():
    seqm, gfm = _ref_modules()
    rng = np.random.default_rng(3)
    b2i = O.base_to_index()
    for _ in range(20):
        s = "".join(np.array(list("ACGTacgtNnUR"))[rng.integers(0, 12, rng.integers(0, 300))].tolist())
        assert np.array_equal(seqm._fast_sequence_to_encoding(s, b2i, 4), O.sequence_to_encoding(s))
    feats = ["f%d" % i for i in range(12)]
    fid = {f: i for i, f in enumerate(feats)}
    g = H.synth_genome(length=30000)
    by = H.rows_by_chrom(H.synth_intervals(g))
    for _ in range(100):
        c = list(g.keys())[rng.integers(0, 3)]
        s = int(rng.integers(0, 29000))
        e = s + int(rng.integers(1, 400))
        thr = rng.random(12).astype(np.float32)
        rows = O.query_rows(by, c, s, e)
        assert np.array_equal(gfm._fast_get_feature_data(s, e, thr.copy(), fid, rows),
                              O.fast_get_feature_data(s, e, thr.copy(), fid, rows))


def test_golden_sampler():
    """RandomPositionsSampler: same RNG streams, same draws, same rejections as the reference."""
    g = H.synth_genome(n_chrom=5, length=6000, n_frac=0.08)
    rows = H.rows_by_chrom(H.synth_intervals(g, n_features=12, per_feature=120, seed=3))
    s = O.RandomPositionsSamplerOracle(g, rows, ["f%d" % i for i in range(12)], seed=436,
                                       validation_holdout=['chr4'], test_holdout=['chr5'])
    a, ta = s.sample(16)
    b, tb = s.sample(5, mode="validate")
    assert np.array_equal(a.astype(np.float32), GOLD["sampler_seq_train"])
    assert np.array_equal(ta.astype(np.float32), GOLD["sampler_tgt_train"])
    assert np.array_equal(b.astype(np.float32), GOLD["sampler_seq_val"])
    assert np.array_equal(tb.astype(np.float32), GOLD["sampler_tgt_val"])
    assert (GOLD["sampler_seq_train"] == 0.25).any() and GOLD["sampler_tgt_train"].sum() > 0


def test_oracle_vep_scores_equal_installed_reference(tmp_path):
    """The oracle port against the UNMODIFIED reference installed in baseline/_ref (baseline/build_ref.py), driven
    through its public AnalyzeSequences.variant_effect_prediction: abs_diffs and logits are bit-identical
    (same torch CPU kernels, same order of operations).  Skipped where the reference is not installed."""
    import bench as B
    from baseline import reference_arm as R
    if not R.available():
        pytest.skip("baseline/_ref not installed (python baseline/build_ref.py in the authoring container)")
    names, seqs = B.synth_genome_arrays(1.2)
    chrom, pos, ref, alt = B.synth_snvs(names, seqs, 70, 5)
    rv = R.ReferenceVEP(names, seqs, str(tmp_path), threads=4)
    got = rv.score(chrom, pos, ref, alt)
    genome = {n: B._SeqView(s) for n, s in zip(names, seqs)}
    variants = [(names[chrom[i]], int(pos[i]), "ACGT"[ref[i]], "ACGT"[alt[i]], "+") for i in range(70)]
    absd, logits, _m, _u = O.vep_scores(genome, variants, O.deepsea_layers(1000, 919, seed=0), 1000)
    assert np.array_equal(got["abs_diffs"], absd)
    assert np.array_equal(got["logits"], logits)
