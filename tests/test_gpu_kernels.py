"""GPU parity tests: every kernel is called through the C ABI (via the ctypes host classes) and
compared with the CPU oracle on identical seeded inputs.  Integer / one-hot / label outputs must be
bit-exact; fp32 predictions <= 1e-5 abs; bf16 predictions <= 2e-2 abs on sigmoid outputs."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import selene_oracle as O  # noqa: E402
import sb_helpers as H  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


@pytest.fixture(scope="module")
def genome_pair(tmp_path_factory):
    from selene_b200.sequences import Genome
    g = H.synth_genome()
    path = str(tmp_path_factory.mktemp("g") / "synth.fa")
    H.write_fasta(g, path)
    Genome.update_bases_order(['A', 'C', 'G', 'T'])
    return g, Genome(path)


# ------------------------------------------------------------------------------------------------
# tcgen05 GEMM core
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(128, 240, 64), (128, 16, 128), (256, 256, 256), (300, 480, 2560),
                                   (1000, 920, 928), (77, 100, 72), (4096, 960, 3840)])
def test_tc_gemm_selftest(torch, M, N, K):
    from selene_b200 import _lib
    lib = _lib.load()
    gen = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = (torch.randn((M, K), device="cuda", generator=gen) * 0.5).to(torch.bfloat16)
    b = (torch.randn((N, K), device="cuda", generator=gen) * 0.5).to(torch.bfloat16)
    c = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32)
    _lib.check(lib.sb_tc_gemm_selftest(_lib.ptr(a), _lib.ptr(b), M, N, K, _lib.ptr(c), _lib.stream_ptr()),
               "sb_tc_gemm_selftest")
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t()
    err = (c - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), "max abs err %g" % err


@pytest.mark.parametrize("pair", ["0", "2"])
@pytest.mark.parametrize("M,N,K", [(129, 240, 192), (256, 480, 640), (385, 256, 64), (1000, 920, 928), (2049, 144, 320)])
def test_tc_gemm_single_cta_and_cta_pairs(torch, monkeypatch, pair, M, N, K):
    """The same GEMMs through 128-row tiles (SB_TC_PAIR=0) and through CTA pairs with cta_group::2 MMAs (SB_TC_PAIR=2
    forces 256-row tiles whatever the wave count): ragged last tiles, a second half that lies wholly past the rows,
    N tiles whose halves are 72 / 120 / 128 rows of B."""
    from selene_b200 import _lib
    monkeypatch.setenv("SB_TC_PAIR", pair)
    lib = _lib.load()
    gen = torch.Generator(device="cuda").manual_seed(M * 5 + N * 11 + K)
    a = (torch.randn((M, K), device="cuda", generator=gen) * 0.5).to(torch.bfloat16)
    b = (torch.randn((N, K), device="cuda", generator=gen) * 0.5).to(torch.bfloat16)
    c = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32)
    _lib.check(lib.sb_tc_gemm_selftest(_lib.ptr(a), _lib.ptr(b), M, N, K, _lib.ptr(c), _lib.stream_ptr()),
               "sb_tc_gemm_selftest")
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t()
    err = (c - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), "max abs err %g" % err


# ------------------------------------------------------------------------------------------------
# kernel (a): encode / codes / flags
# ------------------------------------------------------------------------------------------------
def test_encode_small_fasta_goldens(torch, tmp_path):
    """Known answers of the reference's test_genome.py:40-60 (incl. lowercase and unknowns)."""
    from selene_b200.sequences import Genome
    Genome.update_bases_order(['A', 'C', 'G', 'T'])
    got = Genome.sequence_to_encoding("ATACTTGA")
    exp = np.array([[1, 0, 0, 0], [0, 0, 0, 1], [1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 0, 1],
                    [0, 0, 1, 0], [1, 0, 0, 0]], dtype=np.float32)
    assert got.dtype == np.float32 and np.array_equal(got, exp)
    got = Genome.sequence_to_encoding("AnnUAtCa")
    exp = O.sequence_to_encoding("AnnUAtCa")
    assert np.array_equal(got, exp) and np.array_equal(got[1], np.full(4, 0.25, np.float32))
    assert Genome.sequence_to_encoding("").shape == (0, 4)
    assert Genome.encoding_to_sequence(exp) == "ANNNATCA"


@pytest.mark.parametrize("L", [1, 7, 15, 17, 200, 511, 513, 1000, 1537])
def test_encode_windows_vs_oracle(torch, genome_pair, L):
    g, G = genome_pair
    rng = np.random.default_rng(L)
    chroms, starts, strands = [], [], []
    names = list(g.keys())
    for _ in range(64):
        c = names[rng.integers(0, len(names))]
        # include windows hanging over both ends (pad=True semantics)
        s = int(rng.integers(-L, len(g[c])))
        chroms.append(c)
        starts.append(s)
        strands.append("+-"[rng.integers(0, 2)])
    enc, flags = G.encode_windows(chroms, starts, L, strands=strands)
    enc_b, _ = G.encode_windows(chroms, starts, L, strands=strands, dtype="bfloat16")
    codes, flags2 = G.window_codes(chroms, starts, L, strands=strands)
    enc, flags, enc_b = enc.cpu().numpy(), flags.cpu().numpy(), enc_b.float().cpu().numpy()
    codes = codes.cpu().numpy()
    assert np.array_equal(flags, flags2.cpu().numpy())
    for i in range(64):
        seq = O.get_sequence_from_coords(g, chroms[i], starts[i], starts[i] + L, strands[i], pad=True)
        if seq == "":      # start >= len etc.; the host rejects these before the kernel
            continue
        exp = O.sequence_to_encoding(seq)
        assert np.array_equal(enc[i], exp), (chroms[i], starts[i], strands[i])
        assert np.array_equal(enc_b[i], exp)
        assert bool(flags[i] & 1) == ("N" in seq)
        exp_codes = np.array([{"A": 0, "C": 1, "G": 2, "T": 3}.get(ch.upper(), 4) for ch in seq], np.uint8)
        assert np.array_equal(codes[i], exp_codes)


def test_genome_reference_api(torch, genome_pair):
    g, G = genome_pair
    assert G.get_chrs() == sorted(g.keys())
    assert dict(G.get_chr_lens()) == {k: len(v) for k, v in g.items()}
    c = "chr2"
    n = len(g[c])
    for (s, e, strand, pad) in [(0, 50, "+", False), (100, 1100, "-", False), (-5, 20, "+", True),
                                (n - 10, n + 7, "-", True), (-5, 20, "+", False), (n, n + 5, "+", True),
                                (30, 30, "+", False), (10, 40, ".", False)]:
        exp_s = O.get_sequence_from_coords(g, c, s, e, strand, pad)
        assert G.get_sequence_from_coords(c, s, e, strand, pad) == exp_s
        exp, exp_unk = O.get_encoding_from_coords_check_unk(g, c, s, e, strand, pad)
        got, unk = G.get_encoding_from_coords_check_unk(c, s, e, strand, pad)
        assert got.shape == exp.shape and np.array_equal(got, exp) and unk == exp_unk
    assert G.get_encoding_from_coords("chrZ", 0, 10).shape == (0, 4)
    with pytest.raises(ValueError):
        G.get_encoding_from_coords(c, 0, 10, strand="x")
    assert G.coords_in_bounds(c, 0, n) and not G.coords_in_bounds(c, 0, n + 1)


def test_bases_order_is_a_parameter(torch, genome_pair):
    g, G = genome_pair
    from selene_b200.sequences import Genome
    try:
        Genome.update_bases_order(['A', 'G', 'C', 'T'])
        got = G.get_encoding_from_coords("chr1", 60, 260)
        exp = O.sequence_to_encoding(g["chr1"][60:260], ['A', 'G', 'C', 'T'])
        assert np.array_equal(got, exp)
    finally:
        Genome.update_bases_order(['A', 'C', 'G', 'T'])


# ------------------------------------------------------------------------------------------------
# ISM enumeration + expansion
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seq,rng_", [("ATCCG", None), ("ATCCG", (1, 3)), ("ANNTC", None), (None, None)])
def test_ism_enumerate_and_expand(torch, seq, rng_):
    from selene_b200.predict._in_silico_mutagenesis import ism_mutations_device, expand_mutants_device
    from selene_b200.sequences import Genome
    if seq is None:
        r = np.random.default_rng(0)
        seq = "".join(np.array(list("ACGTN"))[r.choice(5, 1000, p=[.24, .24, .24, .24, .04])].tolist())
    s0, e0 = rng_ if rng_ else (0, len(seq))
    exp = O.in_silico_mutagenesis_sequences(seq, 1, start_position=s0, end_position=e0)
    pos, alt, codes = ism_mutations_device(seq, s0, e0)
    got = [[(int(p), "ACGT"[int(a)])] for p, a in zip(pos.cpu().numpy(), alt.cpu().numpy())]
    assert got == exp
    enc = O.sequence_to_encoding(seq)
    out = expand_mutants_device(codes, pos, alt).cpu().numpy()
    assert out.shape == (len(exp), len(seq), 4)
    for m in range(0, len(exp), max(1, len(exp) // 97)):
        assert np.array_equal(out[m], O.mutate_sequence(enc, exp[m]))


def test_ism_golden_ATCCG(torch):
    """selene_sdk/predict/tests/test_model_predict.py:12-31."""
    from selene_b200.predict._in_silico_mutagenesis import in_silico_mutagenesis_sequences
    observed = in_silico_mutagenesis_sequences("ATCCG")
    expected = [[(0, 'C')], [(0, 'G')], [(0, 'T')], [(1, 'A')], [(1, 'C')], [(1, 'G')], [(2, 'A')], [(2, 'G')],
                [(2, 'T')], [(3, 'A')], [(3, 'G')], [(3, 'T')], [(4, 'A')], [(4, 'C')], [(4, 'T')]]
    assert observed == expected


# ------------------------------------------------------------------------------------------------
# SNV ref/alt pairs
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("L", [1000, 201, 513, 16])
def test_snv_pairs_vs_oracle(torch, genome_pair, L):
    g, G = genome_pair
    rng = np.random.default_rng(2024 + L)
    names = list(g.keys())
    sr, er = O.radii(L)
    var_row = O.get_ref_idxs(L, 1)[0]
    variants = []
    for i in range(96):
        c = names[rng.integers(0, len(names))]
        pos = int(rng.integers(1, len(g[c]) + 1))        # 1-based; windows may hang over the ends
        true_ref = g[c][pos - 1]
        ref = true_ref.upper() if rng.random() > 0.15 else "ACGT"[rng.integers(0, 4)]
        if ref not in "ACGT":
            ref = "N" if rng.random() < 0.5 else "A"
        alt = "ACGTN"[rng.integers(0, 5)]
        variants.append((c, pos, ref, alt, "+-"[rng.integers(0, 2)]))
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    res = G.snv_pairs([v[0] for v in variants], [v[1] - sr for v in variants],
                      [code.get(v[2], 4) for v in variants], [code.get(v[3], 4) for v in variants], L, var_row,
                      strands=[v[4] for v in variants], want_codes=True)
    ref_o, alt_o = res["ref"].cpu().numpy(), res["alt"].cpu().numpy()
    match, flags = res["ref_match"].cpu().numpy(), res["flags"].cpu().numpy()
    # the bf16 tensors and the codes-only call (the path bench.py runs) carry the same information
    res_b = G.snv_pairs([v[0] for v in variants], [v[1] - sr for v in variants],
                        [code.get(v[2], 4) for v in variants], [code.get(v[3], 4) for v in variants], L, var_row,
                        strands=[v[4] for v in variants], dtype="bfloat16", want_codes=True)
    assert np.array_equal(res_b["ref"].float().cpu().numpy(), ref_o) and np.array_equal(res_b["alt"].float().cpu().numpy(), alt_o)
    res_c = G.snv_pairs([v[0] for v in variants], [v[1] - sr for v in variants],
                        [code.get(v[2], 4) for v in variants], [code.get(v[3], 4) for v in variants], L, var_row,
                        strands=[v[4] for v in variants], want_onehot=False, want_codes=True)
    codes = res_c["codes"].cpu().numpy()
    assert np.array_equal(codes, res["codes"].cpu().numpy()) and np.array_equal(codes, res_b["codes"].cpu().numpy())
    assert np.array_equal(res_c["ref_match"].cpu().numpy(), match) and np.array_equal(res_c["flags"].cpu().numpy(), flags)
    from_onehot = np.where(ref_o.max(axis=2) == 1, ref_o.argmax(axis=2), 4)
    for i, v in enumerate(variants):
        o_var = L - 1 - var_row if v[4] == "-" else var_row
        keep = np.arange(L) != o_var           # codes are the unsubstituted window with the strand applied
        assert np.array_equal(codes[i][keep], from_onehot[i][keep])
        if match[i]:
            assert codes[i][o_var] == from_onehot[i][o_var]
    for i, (c, pos, ref, alt, strand) in enumerate(variants):
        start, end = pos - sr, pos + er
        # the reference queries without pad (model_predict.py:1061-1064); out-of-bounds variants are
        # filtered by read_vcf_file, so only in-bounds windows are comparable
        if start < 0 or end > len(g[c]):
            continue
        r, a, m, u = O.snv_ref_alt_encodings(g, c, pos, ref, alt, strand, L)
        assert np.array_equal(ref_o[i], r), variants[i]
        assert np.array_equal(alt_o[i], a), variants[i]
        assert bool(match[i]) == m and bool(flags[i] & 1) == u


# ------------------------------------------------------------------------------------------------
# kernel (b): labels
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("thr", [0.5, 0.29, 1.0, {"default": 0.4, "f3": 0.01}, None])
def test_labels_vs_oracle(torch, tmp_path, thr):
    from selene_b200.targets import GenomicFeatures
    g = H.synth_genome(length=30000)
    rows = H.synth_intervals(g)
    path = str(tmp_path / "feat.bed.gz")
    H.write_bed(rows, path)
    feats = ["f%d" % i for i in range(12)]
    GF = GenomicFeatures(path, feats, feature_thresholds=thr)
    by = H.rows_by_chrom(rows)
    fid = {f: i for i, f in enumerate(feats)}
    rng = np.random.default_rng(5)
    names = list(g.keys())
    qc, qs, qe = [], [], []
    for _ in range(300):
        c = names[rng.integers(0, len(names))]
        s = int(rng.integers(0, len(g[c]) - 200))
        qc.append(c); qs.append(s); qe.append(s + 200)
    got = GF.label_bins(qc, qs, qe, out_dtype="int64").cpu().numpy()
    for i in range(300):
        r = O.query_rows(by, qc[i], qs[i], qe[i])
        if thr is None:
            exp = O.feature_data_no_threshold(12, fid, r)
        else:
            exp = O.fast_get_feature_data(qs[i], qe[i], GF._feature_thresholds_vec, fid, r)
        assert np.array_equal(got[i], exp), (qc[i], qs[i])
    one = GF.get_feature_data(qc[0], qs[0], qe[0])
    assert one.dtype == (np.float64 if thr is None else np.int64) and np.array_equal(one, got[0])
    z = GF.get_feature_data("chrNope", 0, 200)
    assert z.dtype == np.float64 and not z.any()


@pytest.mark.parametrize("thr", [0.5, None])
def test_labels_many_intervals_and_long_bins(torch, tmp_path, thr):
    """Thousands of intervals per chromosome (several rounds of the 32-way search), a few very long intervals
    (the running-max search must still find them), an odd feature count, and bins longer than 2^15 bases
    (two-pass wide state) mixed with short ones in one call."""
    from selene_b200.targets import GenomicFeatures
    rng = np.random.default_rng(17)
    chrom_len = {"chrA": 400000, "chrB": 250000, "chrC": 90000}
    n_feat = 13
    feats = ["f%d" % i for i in range(n_feat)]
    rows = []
    for c, L in chrom_len.items():
        for _ in range(L // 40):
            ln = int(np.clip(rng.lognormal(np.log(150), 1.0), 1, 4000))
            s = int(rng.integers(0, L - ln))
            rows.append((c, s, s + ln, feats[rng.integers(0, n_feat)]))
        for _ in range(6):      # long intervals that reach far past their neighbours
            s = int(rng.integers(0, L // 2))
            rows.append((c, s, s + int(rng.integers(L // 8, L // 2)), feats[rng.integers(0, n_feat)]))
    rows.sort(key=lambda r: (r[0], r[1], r[2]))
    path = str(tmp_path / "many.bed.gz")
    H.write_bed(rows, path)
    GF = GenomicFeatures(path, feats, feature_thresholds=thr)
    by = H.rows_by_chrom(rows)
    fid = {f: i for i, f in enumerate(feats)}
    qc, qs, qe = [], [], []
    names = list(chrom_len)
    for i in range(260):
        c = names[rng.integers(0, 3)]
        ln = 200 if i % 10 else int(rng.integers(32768, 70000))
        s = int(rng.integers(0, chrom_len[c] - ln))
        qc.append(c); qs.append(s); qe.append(s + ln)
    if thr is None:
        got = GF.label_bins(qc, qs, qe, out_dtype="int64").cpu().numpy()     # short and long bins in one launch
    else:
        # thresholds need one bin length per call (the integer threshold depends on it)
        got = np.zeros((260, n_feat), dtype=np.int64)
        short = [i for i in range(260) if qe[i] - qs[i] == 200]
        got[short] = GF.label_bins([qc[i] for i in short], [qs[i] for i in short], [qe[i] for i in short],
                                   out_dtype="int64").cpu().numpy()
        for i in range(260):
            if i not in short:
                got[i] = GF.label_bins([qc[i]], [qs[i]], [qe[i]], out_dtype="int64").cpu().numpy()[0]
    for i in range(260):
        r = O.query_rows(by, qc[i], qs[i], qe[i])
        if thr is None:
            exp = O.feature_data_no_threshold(n_feat, fid, r)
        else:
            exp = O.fast_get_feature_data(qs[i], qe[i], GF._feature_thresholds_vec, fid, r)
        assert np.array_equal(got[i], exp), (qc[i], qs[i], qe[i])


def test_labels_many_contigs_and_negative_starts(torch, tmp_path):
    """More than 512 contigs (the per-contig offsets then stay in global memory instead of shared memory), bins that
    start before position 0 and bins past the last interval of a contig."""
    from selene_b200.targets import GenomicFeatures
    rng = np.random.default_rng(23)
    n_feat = 7
    feats = ["f%d" % i for i in range(n_feat)]
    names = ["ctg%04d" % i for i in range(600)]
    rows = []
    for c in names:
        for _ in range(int(rng.integers(0, 12))):
            s = int(rng.integers(0, 3000))
            rows.append((c, s, s + int(rng.integers(1, 500)), feats[rng.integers(0, n_feat)]))
    rows.sort(key=lambda r: (r[0], r[1], r[2]))
    path = str(tmp_path / "contigs.bed.gz")
    H.write_bed(rows, path)
    GF = GenomicFeatures(path, feats, feature_thresholds=0.3)
    by = H.rows_by_chrom(rows)
    fid = {f: i for i, f in enumerate(feats)}
    qc = [names[rng.integers(0, 600)] for _ in range(400)]
    qs = [int(rng.integers(-150, 4000)) for _ in range(400)]
    qe = [s + 200 for s in qs]
    got = GF.label_bins(qc, qs, qe, out_dtype="int64").cpu().numpy()
    for i in range(400):
        r = O.query_rows(by, qc[i], qs[i], qe[i])
        exp = O.fast_get_feature_data(qs[i], qe[i], GF._feature_thresholds_vec, fid, r)
        assert np.array_equal(got[i], exp), (qc[i], qs[i], qe[i])


def test_labels_reference_goldens(torch, tmp_path):
    """Rows and expected vectors of selene_sdk/targets/tests/test_genomic_features.py:24-42,151-187."""
    from selene_b200.targets import GenomicFeatures
    feats = ["CTCF", "eGFP-FOS", "GABP", "Pbx3", "Pol2", "TBP"]
    rows = [("1", 0, 5, "Pol2"), ("1", 7, 8, "CTCF"), ("1", 14, 18, "TBP"), ("1", 150, 152, "GABP"),
            ("1", 5, 1000, "CTCF"), ("1", 152, 200, "Pbx3"), ("1", 195, 400, "eGFP-FOS")]
    # plus the fixture rows of the reference test module
    rows2 = [("1", 0, 5, "CTCF"), ("1", 4, 8, "eGFP-FOS"), ("1", 0, 10, "GABP"), ("1", 5, 10, "Pbx3"),
             ("1", 5, 6, "Pol2"), ("1", 7, 15, "TBP")]
    for rr, queries in ((rows, [(0, 200), (150, 350)]), (rows2, [(0, 5), (5, 10), (0, 10), (2, 8)])):
        path = str(tmp_path / ("g%d.bed" % len(rr)))
        H.write_bed(sorted(rr, key=lambda r: r[1]), path)
        for thr in (0.5, 0.1, 1.0):
            GF = GenomicFeatures(path, feats, feature_thresholds=thr)
            by = H.rows_by_chrom(rr)
            fid = {f: i for i, f in enumerate(feats)}
            for s, e in queries:
                exp = O.fast_get_feature_data(s, e, GF._feature_thresholds_vec, fid, O.query_rows(by, "1", s, e))
                assert np.array_equal(GF.get_feature_data("1", s, e), exp)


# ------------------------------------------------------------------------------------------------
# kernels (c)+(d): DeepSEA forward + scores
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def deepsea(torch):
    from selene_b200.nets import B200Net
    layers = O.deepsea_layers(1000, 919, seed=0)
    return layers, B200Net(layers, 1000)


def _rand_codes(n, L, seed, p_unk=0.01):
    r = np.random.default_rng(seed)
    return r.choice(5, (n, L), p=[(1 - p_unk) / 4] * 4 + [p_unk]).astype(np.uint8)


def _onehot(codes):
    out = np.zeros(codes.shape + (4,), np.float32)
    for b in range(4):
        out[codes == b, b] = 1
    out[codes == 4] = 0.25
    return out


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_deepsea_forward_vs_oracle(torch, deepsea, precision, tol):
    layers, net = deepsea
    codes = _rand_codes(6, 1000, 11)
    exp = O.stack_forward(_onehot(codes), layers)
    got = net.forward_codes(torch.from_numpy(codes).cuda(), precision=precision).cpu().numpy()
    assert got.shape == (6, 919)
    err = np.abs(got - exp).max()
    assert err <= tol, "max abs err %g" % err
    got2 = net.forward_onehot(torch.from_numpy(_onehot(codes)).cuda(), precision=precision).cpu().numpy()
    assert np.abs(got2 - exp).max() <= tol


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_deepsea_scaled_weights_saturation(torch, precision, tol):
    """up-scaled conv weights exercise saturation and large diffs."""
    from selene_b200.nets import B200Net
    layers = O.deepsea_layers(1000, 919, seed=3, scale=3.0)
    net = B200Net(layers, 1000)
    codes = _rand_codes(4, 1000, 12)
    exp = O.stack_forward(_onehot(codes), layers)
    got = net.forward_codes(torch.from_numpy(codes).cuda(), precision=precision).cpu().numpy()
    assert exp.std() > 0.05  # the scaling really spreads the outputs
    assert np.abs(got - exp).max() <= tol
    # a structural bug (tap order, channel permutation, pooling window) would leave the outputs near
    # 0.5 and still pass an absolute gate at default init; relative to the signal it cannot hide
    assert np.abs(got - exp).max() <= 0.12 * exp.std()
    assert np.corrcoef(got.ravel(), exp.ravel())[0, 1] > 0.999


def test_forward_substitution_rows(torch, deepsea):
    layers, net = deepsea
    codes = _rand_codes(2, 1000, 13)
    src = np.array([0, 0, 1, 1, 0], np.int32)
    pos = np.array([-1, 499, 0, 999, 7], np.int32)
    sub = np.array([0, 2, 4, 1, 3], np.uint8)
    rows = codes[src].copy()
    for i in range(5):
        if pos[i] >= 0:
            rows[i, pos[i]] = sub[i]
    exp = O.stack_forward(_onehot(rows), layers)
    t = lambda a: torch.from_numpy(a).cuda()
    got = net.forward_codes(t(codes), t(src), t(pos), t(sub), n=5, precision="fp32").cpu().numpy()
    assert np.abs(got - exp).max() <= 1e-5


def test_fused_scores_interleaved_and_baseline(torch, deepsea):
    from selene_b200 import _lib
    layers, net = deepsea
    codes = _rand_codes(3, 1000, 14)
    t = lambda a: torch.from_numpy(a).cuda()
    # interleaved ref/alt: rows (ref0, alt0, ref1, alt1, ref2, alt2)
    src = np.repeat(np.arange(3, dtype=np.int32), 2)
    pos = np.full(6, 499, np.int32)
    sub = np.array([codes[0, 499], (codes[0, 499] + 1) % 4, 1, 2, 4, 0], np.uint8)
    rows = codes[src].copy()
    rows[np.arange(6), 499] = sub
    pred = O.stack_forward(_onehot(rows), layers)
    ref, alt = pred[0::2].copy(), pred[1::2].copy()
    res = net.forward_score(t(codes), t(src), t(pos), t(sub), 6, _lib.SB_PAIR_INTERLEAVED, precision="fp32",
                            want=("abs_diffs", "diffs", "logits", "ref_predictions", "alt_predictions"))
    assert np.abs(res["abs_diffs"].cpu().numpy() - O.abs_diff_score(alt, ref)).max() <= 1e-5
    assert np.abs(res["diffs"].cpu().numpy() - O.diff_score(alt, ref)).max() <= 1e-5
    lg = O.logit_score(alt, ref)
    assert np.abs(res["logits"].cpu().numpy() - lg).max() <= 1e-4
    assert np.abs(res["ref_predictions"].cpu().numpy() - ref).max() <= 1e-5
    # baseline mode (ISM): every row scored against one baseline row
    base = t(pred[0:1].copy())
    res = net.forward_score(t(codes), t(src), t(pos), t(sub), 6, _lib.SB_PAIR_BASELINE, precision="fp32",
                            base_pred=base, want=("abs_diffs", "logits"))
    assert np.abs(res["abs_diffs"].cpu().numpy() - np.abs(pred[0:1] - pred)).max() <= 1e-5


def test_score_pairs_clamp_semantics(torch):
    """logit_score_handler.py:118-124: 0 -> 1e-24, >= 1 -> 0.999999, float32 logit."""
    from selene_b200 import _lib
    ref = np.array([[0.0, 1.0, 0.5, 0.25, 1e-30, 0.9999999]], np.float32)
    alt = np.array([[0.5, 0.5, 0.0, 1.0, 0.7, 0.1]], np.float32)
    exp_abs, exp_diff = O.abs_diff_score(alt.copy(), ref.copy()), O.diff_score(alt.copy(), ref.copy())
    exp_logit = O.logit_score(alt.copy(), ref.copy())
    t = lambda a: torch.from_numpy(a).cuda()
    outs = [torch.empty((1, 6), device="cuda") for _ in range(3)]
    ref_d, alt_d = t(ref), t(alt)  # keep the device buffers alive across the call
    _lib.check(_lib.load().sb_score_pairs(_lib.ptr(ref_d), _lib.ptr(alt_d), None, 1, 1, 6, _lib.ptr(outs[0]),
                                          _lib.ptr(outs[1]), _lib.ptr(outs[2]), 0, _lib.stream_ptr()),
               "sb_score_pairs")
    assert np.array_equal(outs[0].cpu().numpy(), exp_abs)
    assert np.array_equal(outs[1].cpu().numpy(), exp_diff)
    assert np.allclose(outs[2].cpu().numpy(), exp_logit, rtol=2e-6, atol=1e-5)


# ------------------------------------------------------------------------------------------------
# the other conv stacks of the reference: HeartENN (BatchNorm folding, renorm, odd row pitches,
# 60/80/240 channels) and Seqweaver (Conv2d (1,8) kernels, 160/320/480 channels)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("arch", ["heartenn", "seqweaver", "deeperdeepsea"])
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_other_conv_stacks_vs_torch(torch, arch, precision, tol):
    import sb_models as M
    from selene_b200.nets import B200Net, layers_from_module
    torch.manual_seed(5)
    if arch == "heartenn":
        model = M.HeartENN(1000, 90)
        M.randomize_batchnorm(model, 1)
        with torch.no_grad():     # make the renorm of heartenn.py:73-78 bite, then apply it once as the host does
            model.conv_net[0].weight.mul_(4.0)
            for m in model.modules():
                if isinstance(m, (torch.nn.Conv1d, torch.nn.Linear)):
                    m.weight.data.renorm_(2, 0, 0.9)
    elif arch == "deeperdeepsea":   # utils/example_model.py: the paper's VEP model
        model = M.DeeperDeepSEA(1000, 919)
        M.randomize_batchnorm(model, 2)
    else:
        model = M.Seqweaver(217)
    model.eval()
    codes = _rand_codes(5, 1000, 21)
    x = _onehot(codes)
    with torch.no_grad():
        exp = model(torch.tensor(x).transpose(1, 2)).numpy()
    net = B200Net(layers_from_module(model), 1000)
    got = net.forward_codes(torch.from_numpy(codes).cuda(), precision=precision).cpu().numpy()
    assert got.shape == exp.shape
    assert np.abs(got - exp).max() <= tol, np.abs(got - exp).max()
    if precision == "bf16":
        assert np.corrcoef(got.ravel(), exp.ravel())[0, 1] > 0.99


# ------------------------------------------------------------------------------------------------
# first conv with many taps and a wide pool (sb_conv1l.cu; DanQ's 26 taps / pool 13 and other shapes)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("taps,pool,cout,L", [(26, 13, 320, 1000), (11, 5, 70, 203), (32, 2, 64, 180), (9, 15, 130, 400),
                                              (12, 3, 8, 97), (10, 7, 257, 300)])
def test_long_first_conv_wide_pool_vs_torch(torch, taps, pool, cout, L):
    import torch.nn as nn
    from selene_b200.nets import B200Net, layers_from_module
    torch.manual_seed(taps * 100 + pool)
    t_pool = (L - taps + 1) // pool

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv_net = nn.Sequential(nn.Conv1d(4, cout, taps), nn.ReLU(inplace=True), nn.MaxPool1d(pool, pool))
            self.classifier = nn.Sequential(nn.Linear(cout * t_pool, 24), nn.Sigmoid())

        def forward(self, x):
            o = self.conv_net(x)
            return self.classifier(o.reshape(o.size(0), -1))
    model = Net().eval()
    with torch.no_grad():
        model.classifier[0].weight.mul_(0.2)
    codes = _rand_codes(9, L, taps + pool, p_unk=0.03)
    x = torch.tensor(_onehot(codes)).transpose(1, 2)
    with torch.no_grad():
        exp = model(x).numpy()
    net = B200Net(layers_from_module(model), L)
    dev_codes = torch.from_numpy(codes).cuda()
    got32 = net.forward_codes(dev_codes, precision="fp32").cpu().numpy()
    assert np.abs(got32 - exp).max() <= 1e-5
    got = net.forward_codes(dev_codes, precision="bf16").cpu().numpy()
    assert np.abs(got - exp).max() <= 2e-2, np.abs(got - exp).max()
    # substituted rows take the same path
    src = torch.tensor([0, 1, 2, 3], dtype=torch.int32).cuda()
    pos = torch.tensor([0, L - 1, L // 2, taps], dtype=torch.int32).cuda()
    sub = torch.tensor([3, 0, 4, 1], dtype=torch.uint8).cuda()
    rows = codes[:4].copy()
    rows[np.arange(4), pos.cpu().numpy()] = sub.cpu().numpy()
    with torch.no_grad():
        exp_rows = model(torch.tensor(_onehot(rows)).transpose(1, 2)).numpy()
    got = net.forward_codes(dev_codes, src, pos, sub, n=4, precision="bf16").cpu().numpy()
    assert np.abs(got - exp_rows).max() <= 2e-2


# ------------------------------------------------------------------------------------------------
# first conv with <= 8 taps and pool 4 on register accumulators (sb_conv1m.cu), odd shapes; and the same net over
# several chunks, where the first conv of chunk k+1 runs on a side stream next to the tcgen05 layers of chunk k
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("taps,cout,L", [(8, 320, 1000), (5, 100, 997), (3, 8, 64), (8, 81, 203), (1, 160, 40), (7, 640, 131)])
def test_pool4_first_conv_register_accumulators_vs_torch(torch, taps, cout, L):
    import torch.nn as nn
    from selene_b200.nets import B200Net, layers_from_module
    torch.manual_seed(taps * 1000 + cout)
    t_pool = (L - taps + 1) // 4

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv_net = nn.Sequential(nn.Conv1d(4, cout, taps), nn.ReLU(inplace=True), nn.MaxPool1d(4, 4))
            self.classifier = nn.Sequential(nn.Linear(cout * t_pool, 24), nn.Sigmoid())

        def forward(self, x):
            o = self.conv_net(x)
            return self.classifier(o.reshape(o.size(0), -1))
    model = Net().eval()
    with torch.no_grad():
        model.classifier[0].weight.mul_(0.2)
    codes = _rand_codes(11, L, taps + cout, p_unk=0.05)
    with torch.no_grad():
        exp = model(torch.tensor(_onehot(codes)).transpose(1, 2)).numpy()
    net = B200Net(layers_from_module(model), L)
    dev_codes = torch.from_numpy(codes).cuda()
    got = net.forward_codes(dev_codes, precision="bf16").cpu().numpy()
    assert np.abs(got - exp).max() <= 2e-2, np.abs(got - exp).max()
    assert np.corrcoef(got.ravel(), exp.ravel())[0, 1] > 0.999
    # substituted rows (first / last / unknown base) through a row map
    src = torch.tensor([0, 1, 2, 3, 3], dtype=torch.int32).cuda()
    pos = torch.tensor([0, L - 1, L // 2, min(taps, L - 1), 7 % L], dtype=torch.int32).cuda()
    sub = torch.tensor([3, 0, 4, 1, 2], dtype=torch.uint8).cuda()
    rows = codes[[0, 1, 2, 3, 3]].copy()
    rows[np.arange(5), pos.cpu().numpy()] = sub.cpu().numpy()
    with torch.no_grad():
        exp_rows = model(torch.tensor(_onehot(rows)).transpose(1, 2)).numpy()
    got = net.forward_codes(dev_codes, src, pos, sub, n=5, precision="bf16").cpu().numpy()
    assert np.abs(got - exp_rows).max() <= 2e-2


def test_first_conv_one_chunk_ahead_equals_single_stream(torch, deepsea, monkeypatch):
    """More rows than one chunk: the side-stream pipeline (first conv of chunk k+1 during the other layers of chunk k,
    two extra buffers) must give bit-identical predictions to the same kernels run chunk by chunk on one stream."""
    from selene_b200.nets import B200Net
    monkeypatch.setenv("SB_CHUNK", "512")
    layers, _ = deepsea
    net = B200Net(layers, 1000)
    n = 512 * 3 + 77
    codes = torch.from_numpy(_rand_codes(64, 1000, 5)).cuda()
    src = torch.from_numpy(np.random.default_rng(1).integers(0, 64, n).astype(np.int32)).cuda()
    pos = torch.from_numpy(np.random.default_rng(2).integers(0, 1000, n).astype(np.int32)).cuda()
    sub = torch.from_numpy(np.random.default_rng(3).integers(0, 5, n).astype(np.uint8)).cuda()
    ahead = net.forward_codes(codes, src, pos, sub, n=n, precision="bf16").cpu().numpy()
    again = net.forward_codes(codes, src, pos, sub, n=n, precision="bf16").cpu().numpy()
    monkeypatch.setenv("SB_CONV1_AHEAD", "0")
    plain = net.forward_codes(codes, src, pos, sub, n=n, precision="bf16").cpu().numpy()
    assert np.array_equal(ahead, again)
    assert np.array_equal(ahead, plain)


# ------------------------------------------------------------------------------------------------
# NonStrandSpecific wrapper: both strands through the network, mean / max of the outputs
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["mean", "max"])
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_non_strand_specific_vs_torch(torch, mode, precision, tol):
    import sb_models as M
    from selene_b200 import _lib
    from selene_b200.nets import B200Net, layers_from_module, unwrap_non_strand_specific
    torch.manual_seed(13)
    wrapped = M.NonStrandSpecific(M.DeepSEA(1000, 919), mode=mode).eval()
    inner, sm = unwrap_non_strand_specific(wrapped)
    assert inner is wrapped.model and sm == (1 if mode == "mean" else 2)
    net = B200Net(layers_from_module(wrapped), 1000, strand_mode=sm)
    codes = _rand_codes(6, 1000, 41)
    x = torch.tensor(_onehot(codes)).transpose(1, 2)
    with torch.no_grad():
        exp = wrapped(x).numpy()
        plain = wrapped.model(x).numpy()
    assert np.abs(exp - plain).max() > 1e-4                     # the wrapper changes the answer
    dev_codes = torch.from_numpy(codes).cuda()
    got = net.forward_codes(dev_codes, precision=precision).cpu().numpy()
    assert np.abs(got - exp).max() <= tol, np.abs(got - exp).max()
    # rows built by substitution (ISM / VEP rows): row i = sequence i % 3 with one base replaced
    src = torch.tensor([0, 1, 2, 0, 1, 2], dtype=torch.int32).cuda()
    pos = torch.tensor([0, 499, 999, 3, 500, 7], dtype=torch.int32).cuda()
    sub = torch.tensor([1, 2, 3, 0, 4, 2], dtype=torch.uint8).cuda()
    rows = codes[[0, 1, 2, 0, 1, 2]].copy()
    rows[np.arange(6), pos.cpu().numpy()] = sub.cpu().numpy()
    with torch.no_grad():
        exp_rows = wrapped(torch.tensor(_onehot(rows)).transpose(1, 2)).numpy()
    got = net.forward_codes(dev_codes, src, pos, sub, n=6, precision=precision).cpu().numpy()
    assert np.abs(got - exp_rows).max() <= tol
    # interleaved ref/alt scoring on the combined predictions
    res = net.forward_score(dev_codes, src, pos, sub, 6, _lib.SB_PAIR_INTERLEAVED, precision=precision,
                            want=("abs_diffs", "diffs", "logits"))
    ref_p, alt_p = exp_rows[0::2], exp_rows[1::2]
    assert np.abs(res["abs_diffs"].cpu().numpy() - np.abs(ref_p - alt_p)).max() <= 2 * tol
    assert np.abs(res["diffs"].cpu().numpy() - (alt_p - ref_p)).max() <= 2 * tol
    if precision == "fp32":
        # arbitrary float input (not one-hot): the generic path flips both axes
        xg = torch.rand((3, 1000, 4), generator=torch.Generator().manual_seed(3))
        with torch.no_grad():
            exp_g = wrapped(xg.transpose(1, 2)).numpy()
        got_g = net.forward_onehot(xg.cuda(), precision="fp32").cpu().numpy()
        assert np.abs(got_g - exp_g).max() <= tol
        # another channel order (A, G, C, T): the wrapper flips CHANNELS, whatever bases they hold
        perm = (0, 2, 1, 3)
        net2 = B200Net(layers_from_module(wrapped), 1000, perm=perm, strand_mode=mode)
        oh = np.zeros(codes.shape + (4,), np.float32)
        for b in range(4):
            oh[codes == b, perm[b]] = 1
        oh[codes == 4] = 0.25
        with torch.no_grad():
            exp2 = wrapped(torch.tensor(oh).transpose(1, 2)).numpy()
        got2 = net2.forward_codes(dev_codes, precision="fp32").cpu().numpy()
        assert np.abs(got2 - exp2).max() <= tol


# ------------------------------------------------------------------------------------------------
# DanQ: conv(k=26)+pool13 -> BiLSTM whose recurrence runs over the batch rows (reference quirk) -> FC
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_danq_batch_axis_recurrence_vs_torch(torch, precision, tol):
    import sb_models as M
    from selene_b200.nets import B200Net, layers_from_module
    torch.manual_seed(9)
    model = M.DanQ(1000, 919).eval()
    codes = _rand_codes(7, 1000, 31)
    x = torch.tensor(_onehot(codes)).transpose(1, 2)
    net = B200Net(layers_from_module(model), 1000)
    dev_codes = torch.from_numpy(codes).cuda()
    with torch.no_grad():
        whole = model(x).numpy()                                       # one predict() call of 7 rows
        split = np.concatenate([model(x[:3]).numpy(), model(x[3:6]).numpy(), model(x[6:]).numpy()])
    assert np.abs(whole - split).max() > 1e-4                          # the quirk is real: grouping matters
    net.set_recurrence(7, 1)
    got = net.forward_codes(dev_codes, precision=precision).cpu().numpy()
    assert np.abs(got - whole).max() <= tol, np.abs(got - whole).max()
    net.set_recurrence(3, 1)
    got = net.forward_codes(dev_codes, precision=precision).cpu().numpy()
    assert np.abs(got - split).max() <= tol, np.abs(got - split).max()
    if precision == "fp32":
        # interleaved ref/alt rows: the even rows form one chain, the odd rows another
        net.set_recurrence(4, 2)
        got = net.forward_codes(dev_codes[:6], precision=precision).cpu().numpy()
        with torch.no_grad():
            exp_even, exp_odd = model(x[0:6:2]).numpy(), model(x[1:6:2]).numpy()
        assert np.abs(got[0::2] - exp_even).max() <= tol and np.abs(got[1::2] - exp_odd).max() <= tol


@pytest.mark.parametrize("stepwise,cluster", [("0", "1"), ("0", "0"), ("1", "0")])
def test_danq_recurrence_paths(torch, monkeypatch, stepwise, cluster):
    """bf16 recurrence three ways — the one-launch persistent kernel (sb_lstm.cu) with 2-CTA clusters exchanging h
    through distributed shared memory (default for small jobs), the same kernel with one CTA per unit tile, and
    the per-step launches (SB_LSTM_STEPWISE=1) — against torch, ragged groups and interleaved ref/alt chains
    included."""
    import sb_models as M
    from selene_b200.nets import B200Net, layers_from_module
    monkeypatch.setenv("SB_LSTM_STEPWISE", stepwise)
    monkeypatch.setenv("SB_LSTM_CLUSTER", cluster)
    torch.manual_seed(10)
    model = M.DanQ(1000, 919).eval()
    codes = _rand_codes(11, 1000, 32)
    x = torch.tensor(_onehot(codes)).transpose(1, 2)
    net = B200Net(layers_from_module(model), 1000)
    with torch.no_grad():
        split = np.concatenate([model(x[:4]).numpy(), model(x[4:8]).numpy(), model(x[8:]).numpy()])
        exp_even, exp_odd = model(x[0:10:2]).numpy(), model(x[1:10:2]).numpy()
    net.set_recurrence(4, 1)
    got = net.forward_codes(torch.from_numpy(codes).cuda(), precision="bf16").cpu().numpy()
    assert np.abs(got - split).max() <= 2e-2, np.abs(got - split).max()
    net.set_recurrence(5, 2)
    got = net.forward_codes(torch.from_numpy(codes[:10]).cuda(), precision="bf16").cpu().numpy()
    assert np.abs(got[0::2] - exp_even).max() <= 2e-2 and np.abs(got[1::2] - exp_odd).max() <= 2e-2


# ------------------------------------------------------------------------------------------------
# MN-major split-K wgrad GEMM (training)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,M,x_ld,N", [(64, 128, 256, 256), (200, 64, 64, 64), (1000, 480, 320, 2560),
                                           (333, 960, 480, 3840), (64, 920, 1024, 1024), (4000, 96, 64, 512)])
def test_tc_wgrad_selftest(torch, rows, M, x_ld, N):
    from selene_b200 import _lib
    lib = _lib.load()
    gen = torch.Generator(device="cuda").manual_seed(rows + M + N)
    dy = (torch.randn((rows, M), device="cuda", generator=gen) * 0.5).to(torch.bfloat16)
    x = (torch.randn((rows, x_ld), device="cuda", generator=gen) * 0.5).to(torch.bfloat16)
    g = torch.zeros((N, M), device="cuda", dtype=torch.float32)
    _lib.check(lib.sb_tc_wgrad_selftest(_lib.ptr(dy), _lib.ptr(x), rows, M, x_ld, N, _lib.ptr(g), _lib.stream_ptr()),
               "sb_tc_wgrad_selftest")
    torch.cuda.synchronize()
    # row r of the view = flat x[r*x_ld : r*x_ld + N], zero beyond the buffer / for rows whose view overruns
    span = (N + x_ld - 1) // x_ld
    flat = torch.cat([x.float().reshape(-1), torch.zeros(N, device="cuda")])
    view = torch.stack([flat[r * x_ld: r * x_ld + N] for r in range(rows)])
    if span > 1:
        view[rows - (span - 1):] = 0
    ref = view.t() @ dy.float()
    err = (g - ref).abs().max().item()
    assert err <= 3e-3 * max(1.0, ref.abs().max().item()), "max abs err %g" % err


# ------------------------------------------------------------------------------------------------
# the writers' "{:.2e}" formatting on the device
# ------------------------------------------------------------------------------------------------
def _py_rows(a):
    return ["\t".join("{:.2e}".format(v) for v in row) + "\n" for row in a]


def test_format_e2_matches_python(torch):
    from selene_b200.predict.model_predict import format_e2_rows
    special = np.array([0.0, -0.0, 1.0, -1.0, 0.125, 1.125, 1.375, 1.135, 2.5, 9.995, 9.9951, 99.95, 999.5, 0.5, 0.999999,
                        1e-24, 123.456, 1.005, 1.015, 1.025, 3.4028235e38, -3.4028235e38, 1.4e-45, 1.1754944e-38,
                        np.nan, np.inf, -np.inf, 6.5e-5, 1.0000001, 0.99999994, 9.99e10, 1.245e-3, 4.4999999e-7,
                        1e10, 1.5e-10, 2.675, 1e-3, 1.0625, 1.1875, 8.5e8, 1.00097656], dtype=np.float32)
    r = np.random.default_rng(0)
    rand_bits = r.integers(0, 2 ** 32, (2048, 257), dtype=np.uint64).astype(np.uint32).view(np.float32)
    # exact ties on purpose: k + 0.5 at three significant digits, scaled by powers of ten / two that stay exact
    ties = np.array([(2 * k + 1) / 2.0 * s for k in range(100, 1000, 7) for s in (1e-2, 1e-1, 1.0, 10.0, 1e3, 1e5)],
                    dtype=np.float32)
    small = (r.random((512, 919)) * np.float32(10.0) ** r.integers(-8, 1, (512, 919))).astype(np.float32)
    for a in (special[None, :], rand_bits, ties[None, :], small, np.zeros((3, 1), np.float32)):
        buf, off = format_e2_rows(a)
        got = [bytes(buf[off[i]:off[i + 1]]).decode() for i in range(a.shape[0])]
        exp = _py_rows(a)
        bad = [i for i in range(len(exp)) if got[i] != exp[i]]
        if bad:
            g, e = got[bad[0]].split("\t"), exp[bad[0]].split("\t")
            diff = [(a[bad[0], j], g[j], e[j]) for j in range(len(e)) if j >= len(g) or g[j] != e[j]][:5]
            raise AssertionError("row %d differs: %r" % (bad[0], diff))
    buf, off = format_e2_rows(np.zeros((0, 5), np.float32))
    assert buf.size == 0 and off.tolist() == [0]
    # id columns spliced in on the device (sb_format_e2_rows_prefixed), empty and long prefixes included
    pre = [b"", b"12\tA\tC\t", b"x" * 70 + b"\t", b"chr1\t1000\trs1\tA\tG\t"] * 8
    a = small[:32, :40]
    buf, off = format_e2_rows(a, pre)
    exp = _py_rows(a)
    for i in range(32):
        assert bytes(buf[off[i]:off[i + 1]]) == pre[i] + exp[i].encode()
    assert bytes(buf) == b"".join(p + e.encode() for p, e in zip(pre, exp))


def test_tsv_writer_uses_device_formatter(torch, tmp_path):
    """handler.py:15-42,99-114: header, id columns, then "{:.2e}" per value; chunks are written concurrently at their
    byte offsets (several 32 MB chunks for the large matrix), from a producer stream other than the default one."""
    from selene_b200.predict.model_predict import write_tsv
    from selene_b200.predict import _writers
    data = np.array([[1.234e-5, 0.5, 0.999999], [1e-24, 123.456, 0.0]], dtype=np.float32)
    path = str(tmp_path / "x.tsv")
    ids = [("1", "A", "C"), ("2", "G", "T")]
    write_tsv(path, ["pos", "ref", "alt"], ["f0", "f1", "f2"], ids, data)
    lines = open(path).read().split("\n")
    assert lines[0] == "pos\tref\talt\tf0\tf1\tf2" and lines[-1] == "" and len(lines) == 4
    for line, row, ident in zip(lines[1:3], [data[0], data[1]], ids):
        assert line == "\t".join(ident) + "\t" + "\t".join("{:.2e}".format(p) for p in list(row))
    write_tsv(path, ["pos"], ["f0"], [], np.zeros((0, 1), np.float32))          # no rows: header only
    assert open(path).read() == "pos\tf0\n"
    # ~100 MB of text -> 4 chunks on the writer pool; produced on a side stream
    r = np.random.default_rng(1)
    n, F = 12000, 919
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        big = torch.from_numpy((r.random((n, F)) * 10.0 ** r.integers(-6, 1, (n, F))).astype(np.float32)).cuda()
        big = big * 1.0
        ids = [("chr1", i, "rs%d" % i) for i in range(n)]
        finish = _writers.submit_tsv(path, ["chrom", "pos", "id"], ["f%d" % i for i in range(F)], big,
                                     *_writers.prefixes_from_ids(ids))
    finish()
    host = big.cpu().numpy()
    with open(path) as fh:
        assert fh.readline().count("\t") == F + 2
        for i, line in enumerate(fh):
            if i % 997 == 0 or i == n - 1:
                assert line == "chr1\t%d\trs%d\t" % (i, i) + "\t".join("{:.2e}".format(p) for p in list(host[i])) + "\n"
    assert i == n - 1
