"""Vectorised VCF ingest (selene_b200/predict/_variant_effect_prediction.py: VariantTable) against the reference's
``read_vcf_file`` semantics (selene_sdk/predict/_variant_effect_prediction.py:13-134): a hand-checked expectation
that always runs, and a differential run against the installed unmodified reference (baseline/_ref) on files that
exercise every branch."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))


class FakeGenome(object):
    """The part of Genome that the VCF reader touches."""
    _blacklist = None

    def __init__(self, lens):
        self.len_chrs = dict(lens)

    def get_chrs(self):
        return list(self.len_chrs.keys())

    def coords_in_bounds(self, chrom, start, end):
        n = self.len_chrs.get(chrom)
        return n is not None and start < n and start < end and end > 0 and start >= 0 and end <= n


BODY = [
    "chr1\t600\trs1\tA\tG",
    "chr1\t601\trs.2\tA\tG,T",                 # two alts
    "chr1\t602\t.\tAC\tA.T",                   # '.' separates alts too
    "1\t700\tnochr\tC\tT\textra\tcols",        # chr prefix added
    "CHR2\t800\tupper\tG\t-",                  # CHR -> chr, deletion alt
    "chr2\t900\tins\t-\tACG",                  # '-' ref = empty
    "chr2\t950\tstar\tT\t*",
    "chrMT\t1000\tmito\tA\tC",                 # chrMT -> chrM (genome has chrM)
    "chr1\t100\toob_left\tA\tC",               # window out of bounds -> NA
    "chr1\t4900\toob_right\tA\tC",
    "chr9\t1000\tunknown_chrom\tA\tC",         # not in the genome -> NA
    "chr1\t700\tfew\tA",                       # fewer than 5 columns -> NA
    "",                                        # blank line -> NA
    "chr1\t0750\tzeros\tA\tC",                 # int('0750') == 750
    "  chr1\t760\tws\tA\tC  ",                 # strip()
    "chr1\t770\tdot_alt\tA\t.",                # '.' alone: two empty alternatives
    "chr1\t780\tlong\tACGTACGT\tA",
    "chr2\t1500\tn_ref\tN\ta",
]


def _write(path, lines, newline="\n", header=True, trailing=True):
    head = ["##fileformat=VCFv4.2", "##source=test", "#CHROM\tPOS\tID\tREF\tALT\tSTRAND"] if header else []
    text = newline.join(head + lines) + (newline if trailing else "")
    with open(path, "wb") as fh:
        fh.write(text.encode())


GENOME = {"chr1": 5000, "chr2": 4000, "chrM": 3000}


def test_vcf_table_hand_checked(tmp_path):
    from selene_b200.predict._variant_effect_prediction import read_vcf_file, VariantTable
    p = str(tmp_path / "a.vcf")
    _write(p, BODY)
    na = str(tmp_path / "a.NA")
    got = read_vcf_file(p, output_NAs_to_file=na, seq_context=(500, 500), reference_sequence=FakeGenome(GENOME))
    exp = [("chr1", 600, "rs1", "A", "G", "+"), ("chr1", 601, "rs.2", "A", "G", "+"), ("chr1", 601, "rs.2", "A", "T", "+"),
           ("chr1", 602, ".", "AC", "A", "+"), ("chr1", 602, ".", "AC", "T", "+"), ("chr1", 700, "nochr", "C", "T", "+"),
           ("chr2", 800, "upper", "G", "-", "+"), ("chr2", 900, "ins", "", "ACG", "+"), ("chr2", 950, "star", "T", "*", "+"),
           ("chrM", 1000, "mito", "A", "C", "+"), ("chr1", 750, "zeros", "A", "C", "+"), ("chr1", 760, "ws", "A", "C", "+"),
           ("chr1", 770, "dot_alt", "A", "", "+"), ("chr1", 770, "dot_alt", "A", "", "+"),
           ("chr1", 780, "long", "ACGTACGT", "A", "+"), ("chr2", 1500, "n_ref", "N", "a", "+")]
    assert got == exp
    assert open(na).read() == "".join(l + "\n" for l in (BODY[8], BODY[9], BODY[10], BODY[11], BODY[12]))
    t = VariantTable.from_vcf(p, seq_context=(500, 500), reference_sequence=FakeGenome(GENOME))
    snv = [i for i, v in enumerate(exp) if len(v[3]) == 1 and len(v[4]) == 1 and v[4] not in "*-"]
    assert np.flatnonzero(t.is_snv).tolist() == snv
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    assert [int(t.ref_code[i]) for i in snv] == [code.get(exp[i][3].upper(), 4) for i in snv]
    assert [int(t.alt_code[i]) for i in snv] == [code.get(exp[i][4].upper(), 4) for i in snv]
    # id columns exactly as handler.py:36-42 writes them
    match = np.array([i % 3 != 0 for i in range(len(exp))])
    unk = np.array([i % 4 == 1 for i in range(len(exp))])
    pre, off = t.id_prefix(match, unk)
    for i, v in enumerate(exp):
        want = ("\t".join(str(x) for x in v + (bool(match[i]), bool(unk[i]))) + "\t").encode()
        assert bytes(pre[off[i]:off[i + 1]]) == want
    sub_pre, sub_off = t.rows(3, 9).id_prefix(match[3:9], unk[3:9])
    assert bytes(sub_pre) == bytes(pre[off[3]:off[9]])


def test_vcf_header_rules_and_errors(tmp_path):
    from selene_b200.predict._variant_effect_prediction import read_vcf_file
    G = FakeGenome(GENOME)
    p = str(tmp_path / "h.vcf")
    with open(p, "w") as fh:
        fh.write("#CHROM\tPOS\tID\tALT\tREF\nchr1\t600\tx\tA\tC\n")
    with pytest.raises(ValueError, match="First 5 columns"):
        read_vcf_file(p, seq_context=(500, 500), reference_sequence=G)
    with open(p, "w") as fh:
        fh.write("chr1\tabc\tx\tA\tC\n")
    with pytest.raises(ValueError):
        read_vcf_file(p, seq_context=(500, 500), reference_sequence=G)
    with open(p, "w") as fh:
        fh.write("chr1\t600\tx\tA\tC\n")
    with pytest.raises(IndexError):
        read_vcf_file(p, strand_index=7, seq_context=(500, 500), reference_sequence=G)
    open(p, "w").close()
    assert read_vcf_file(p, seq_context=(500, 500), reference_sequence=G) == []


CASES = [dict(), dict(newline="\r\n"), dict(header=False), dict(trailing=False), dict(strand_index=5),
         dict(strand_index=5, require_strand=True), dict(strand_index=-1), dict(seq_context=None),
         dict(genome={"1": 5000, "2": 4000, "MT": 3000}), dict(genome={"chr1": 5000, "chr2": 4000, "chrMT": 3000})]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_vcf_table_equals_installed_reference(tmp_path, case):
    from baseline import reference_arm as R
    if not R.available():
        pytest.skip("baseline/_ref not installed")
    for q in (R.SHIMS, R.REF_DIR):
        if q not in sys.path:
            sys.path.insert(0, q)
    from selene_sdk.predict._variant_effect_prediction import read_vcf_file as ref_read
    from selene_b200.predict._variant_effect_prediction import read_vcf_file
    kw = dict(CASES[case])
    rng = np.random.default_rng(case)
    strands = ["+", "-", "."]
    lines = []
    for l in BODY:
        if l.count("\t") >= 4 and not l.startswith(" ") and "extra" not in l:
            l = l + "\t" + strands[rng.integers(0, 3)]
        lines.append(l)
    if kw.get("strand_index") is not None:
        lines = [l for l in lines if l.count("\t") >= 5]         # upstream raises IndexError on short rows
    # a few thousand random rows on top of the hand-written ones
    for i in range(3000):
        c = ["chr1", "chr2", "2", "CHR1", "chrM", "chr7"][rng.integers(0, 6)]
        ref = "".join("ACGTN"[k] for k in rng.integers(0, 5, rng.integers(1, 4))) if rng.random() < 0.2 else "ACGT"[rng.integers(0, 4)]
        alt = ",".join("".join("ACGT"[k] for k in rng.integers(0, 4, rng.integers(1, 3))) for _ in range(rng.integers(1, 3)))
        lines.append("%s\t%d\tid%d\t%s\t%s\t%s" % (c, rng.integers(1, 5200), i, ref, alt, strands[rng.integers(0, 3)]))
    p = str(tmp_path / "c.vcf")
    _write(p, lines, newline=kw.pop("newline", "\n"), header=kw.pop("header", True), trailing=kw.pop("trailing", True))
    G = FakeGenome(kw.pop("genome", GENOME))
    ctx = kw.pop("seq_context", (500, 500))
    exp = ref_read(p, output_NAs_to_file=str(tmp_path / "ref.NA"), seq_context=ctx, reference_sequence=G, **kw)
    got = read_vcf_file(p, output_NAs_to_file=str(tmp_path / "got.NA"), seq_context=ctx, reference_sequence=G, **kw)
    assert got == exp
    if ctx:
        assert open(str(tmp_path / "got.NA")).read() == open(str(tmp_path / "ref.NA")).read()
