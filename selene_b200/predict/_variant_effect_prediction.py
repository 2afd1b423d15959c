"""Variant ingest for variant effect prediction — SURVEY §8 f2.

``VariantTable.from_vcf`` parses a VCF file COLUMNAR and VECTORISED (numpy over the byte buffer: line and tab
positions, field slices, integer parse, chromosome normalisation on the handful of distinct names, bounds /
blacklist checks, multi-alt expansion) and keeps the variants as arrays: chromosome index, position, field slices
into the file buffer, strand, SNV codes.  The SNV bulk goes to ``sb_snv_pairs`` straight from these arrays and the
id columns of the output files are assembled from the same slices (``id_prefix``), so no per-variant Python object
is created on the scoring path.  ``read_vcf_file`` — the reference's function, same arguments, same return value
(list of ``(chrom, pos, id, ref, alt, strand)``), same ``.NA`` file — is a thin view of the table.

Semantics follow selene_sdk/predict/_variant_effect_prediction.py:13-134 (header rule, ``strip().split('\\t')``,
CHR/chr/MT renaming, ``-`` ref, strand column, ``require_strand``, the ``pos + len(ref)//2 -/+ radius`` bounds check,
``.`` and ``,`` both splitting ALT).  Variants other than SNVs get their ref/alt windows built string-wise exactly
as ``_process_alt`` / ``_handle_standard_ref`` / ``_handle_long_ref`` do (:146-266) and are scored from code rows.
"""
import math

import numpy as np

VCF_REQUIRED_COLS = ["#CHROM", "POS", "ID", "REF", "ALT"]

_WS = np.zeros(256, dtype=bool)
_WS[[9, 10, 11, 12, 13, 28, 29, 30, 31, 32]] = True            # what str.strip() removes (ASCII range)
_BASE_CODE = np.full(256, 4, dtype=np.uint8)
for _i, _b in enumerate(b"ACGT"):
    _BASE_CODE[_b] = _i
    _BASE_CODE[_b | 0x20] = _i


def _ragged_copy(dst, dst_start, src, src_start, lens):
    """dst[dst_start[i] : +lens[i]] = src[src_start[i] : +lens[i]] for every i, without a Python loop."""
    total = int(lens.sum())
    if total == 0:
        return
    first = np.cumsum(lens) - lens
    within = np.arange(total, dtype=np.int64) - np.repeat(first, lens)
    dst[np.repeat(dst_start, lens) + within] = src[np.repeat(src_start, lens) + within]


def _decimal_text(values):
    """ASCII decimal text of non-negative int64 ``values``: (uint8 buffer, starts, lens)."""
    v = np.asarray(values, dtype=np.int64)
    nd = np.ones(len(v), dtype=np.int64)
    p = 10
    for _ in range(18):
        nd += v >= p
        p *= 10
    starts = np.cumsum(nd) - nd
    buf = np.empty(int(nd.sum()), dtype=np.uint8)
    rest = v.copy()
    for k in range(int(nd.max()) if len(v) else 0):                # digit k from the right
        sel = nd > k
        buf[(starts + nd - 1 - k)[sel]] = 48 + (rest[sel] % 10)
        rest //= 10
    return buf, starts, nd


def _normalise_chrom(chrom, check_chr, genome_chrs):
    """:86-96 of the reference function."""
    if chrom[:3] == 'CHR':
        chrom = chrom.replace('CHR', 'chr')
    elif "chr" not in chrom and check_chr:
        chrom = "chr" + chrom
    if chrom in ("chrMT", "MT") and chrom not in genome_chrs:
        chrom = chrom[:-1]
    return chrom


class VariantTable(object):
    """Columnar variant list.  Field ``x`` of variant ``i`` is ``buf[x_s[i]:x_e[i]]`` (uint8 file bytes)."""

    def __init__(self, buf, chrom_names, chrom_id, pos, name_s, name_e, ref_s, ref_e, alt_s, alt_e, minus):
        self.buf = buf
        self.chrom_names = list(chrom_names)
        self.chrom_id = np.asarray(chrom_id, dtype=np.int32)
        self.pos = np.asarray(pos, dtype=np.int64)
        self.name_s, self.name_e = name_s, name_e
        self.ref_s, self.ref_e = ref_s, ref_e
        self.alt_s, self.alt_e = alt_s, alt_e
        self.minus = np.asarray(minus, dtype=bool)
        ref_len, alt_len = ref_e - ref_s, alt_e - alt_s
        n = len(self.pos)
        first = lambda s, ln: buf[np.minimum(s, max(len(buf) - 1, 0))] if len(buf) else np.zeros(n, np.uint8)
        alt0 = first(alt_s, alt_len)
        # a substitution of one base by one base ('*' and '-' mean "deleted", :180-181)
        self.is_snv = (ref_len == 1) & (alt_len == 1) & (alt0 != 42) & (alt0 != 45)
        self.ref_code = np.where(self.is_snv, _BASE_CODE[first(ref_s, ref_len)], 4).astype(np.uint8)
        self.alt_code = np.where(self.is_snv, _BASE_CODE[alt0], 4).astype(np.uint8)

    def __len__(self):
        return len(self.pos)

    # ------------------------------------------------------------------------------------------ views
    def _text(self, s, e, i):
        return bytes(self.buf[s[i]:e[i]]).decode("utf-8", "replace")

    def variant(self, i):
        """The reference's tuple (chrom, pos, id, ref, alt, strand) of variant ``i``."""
        return (self.chrom_names[self.chrom_id[i]], int(self.pos[i]), self._text(self.name_s, self.name_e, i),
                self._text(self.ref_s, self.ref_e, i), self._text(self.alt_s, self.alt_e, i),
                '-' if self.minus[i] else '+')

    def to_tuples(self):
        return [self.variant(i) for i in range(len(self))]

    def rows(self, lo, hi):
        """Variants [lo, hi) as a table over the same buffer (rank shards)."""
        sl = slice(lo, hi)
        return VariantTable(self.buf, self.chrom_names, self.chrom_id[sl], self.pos[sl], self.name_s[sl],
                            self.name_e[sl], self.ref_s[sl], self.ref_e[sl], self.alt_s[sl], self.alt_e[sl],
                            self.minus[sl])

    def id_prefix(self, ref_match, contains_unk, lo=0, hi=None):
        """Bytes of the id columns of rows [lo, hi) as the TSV writers need them:
        ``chrom \\t pos \\t id \\t ref \\t alt \\t strand \\t ref_match \\t contains_unk \\t`` per row
        (model_predict.py:40, handler.py:36-42 with ``str()`` of every id).  Returns (uint8 buffer, offsets (m+1,))."""
        hi = len(self) if hi is None else hi
        m = hi - lo
        sl = slice(lo, hi)
        cn = [c.encode() for c in self.chrom_names]
        cbuf = np.frombuffer(b"".join(cn), dtype=np.uint8) if cn else np.zeros(0, np.uint8)
        clen = np.array([len(c) for c in cn], dtype=np.int64)
        cstart = np.cumsum(clen) - clen
        cid = self.chrom_id[sl]
        pbuf, pstart, plen = _decimal_text(self.pos[sl])
        tails = [("\t%s\t%s\t%s\t" % (st, mt, uk)).encode() for st in "+-" for mt in ("True", "False")
                 for uk in ("False", "True")]
        tbuf = np.frombuffer(b"".join(tails), dtype=np.uint8)
        tlen = np.array([len(t) for t in tails], dtype=np.int64)
        tstart = np.cumsum(tlen) - tlen
        tid = (self.minus[sl].astype(np.int64) * 4 + (~np.asarray(ref_match[sl], dtype=bool)).astype(np.int64) * 2 +
               np.asarray(contains_unk[sl], dtype=bool).astype(np.int64))
        pieces = [(cbuf, cstart[cid], clen[cid]), (pbuf, pstart, plen),
                  (self.buf, self.name_s[sl], self.name_e[sl] - self.name_s[sl]),
                  (self.buf, self.ref_s[sl], self.ref_e[sl] - self.ref_s[sl]),
                  (self.buf, self.alt_s[sl], self.alt_e[sl] - self.alt_s[sl])]
        row_len = sum(p[2] for p in pieces) + 4 + tlen[tid]      # 4 tabs between the five fields; the tail brings its own
        off = np.zeros(m + 1, dtype=np.int64)
        np.cumsum(row_len, out=off[1:])
        out = np.full(int(off[-1]), 9, dtype=np.uint8)           # tabs everywhere the pieces do not write
        cur = off[:-1].copy()
        for src, s, ln in pieces:
            _ragged_copy(out, cur, src, np.asarray(s, dtype=np.int64), np.asarray(ln, dtype=np.int64))
            cur = cur + ln + 1
        _ragged_copy(out, cur - 1, tbuf, tstart[tid], tlen[tid])
        return out, off

    # ------------------------------------------------------------------------------------------ ingest
    @classmethod
    def from_vcf(cls, input_path, strand_index=None, require_strand=False, output_NAs_to_file=None,
                 seq_context=None, reference_sequence=None):
        with open(input_path, "rb") as fh:
            raw = fh.read()
        buf = np.frombuffer(raw, dtype=np.uint8)
        nb = len(buf)
        genome_chrs = list(reference_sequence.get_chrs())
        check_chr = all(c.startswith("chr") for c in genome_chrs)
        # ---- lines (universal newlines: \n, \r\n, lone \r)
        is_nl = buf == 10
        cr = np.flatnonzero(buf == 13)
        if len(cr):
            lone = cr[(cr + 1 >= nb) | (buf[np.minimum(cr + 1, nb - 1)] != 10)]
            is_nl = is_nl.copy()
            is_nl[lone] = True
        nl = np.flatnonzero(is_nl)
        line_s = np.concatenate([[0], nl + 1]).astype(np.int64)
        line_e = np.concatenate([nl, [nb]]).astype(np.int64)      # exclusive, newline not included
        if len(line_s) and line_s[-1] >= nb:                       # the file ends with a newline: no empty last line
            line_s, line_e = line_s[:-1], line_e[:-1]
        n_lines = len(line_s)
        # ---- header: skip lines up to and including "#CHROM", or up to the first line without '#' (:64-77)
        hashes = np.flatnonzero(buf == 35)
        has_hash = np.searchsorted(hashes, line_e) > np.searchsorted(hashes, line_s)
        first_data = 0
        for i in range(n_lines):
            first_data = i
            if not has_hash[i]:
                break
            text = bytes(buf[line_s[i]:line_e[i]]).decode("utf-8", "replace")
            if "#CHROM" in text:
                cols = text.strip().split('\t')
                if cols[:5] != VCF_REQUIRED_COLS:
                    raise ValueError("First 5 columns in file {0} were {1}. Expected columns: {2}".format(
                        input_path, cols[:5], VCF_REQUIRED_COLS))
                first_data = i + 1
                break
        line_s, line_e = line_s[first_data:], line_e[first_data:]
        n = len(line_s)
        # ---- strip() then split('\t')
        nonws = np.flatnonzero(~_WS[buf])
        lo = np.searchsorted(nonws, line_s)
        hi = np.searchsorted(nonws, line_e)
        blank = hi <= lo
        a = np.where(blank, line_s, nonws[np.minimum(lo, max(len(nonws) - 1, 0))] if len(nonws) else line_s)
        b = np.where(blank, line_s, (nonws[np.maximum(hi - 1, 0)] + 1) if len(nonws) else line_s)
        tabs = np.flatnonzero(buf == 9)
        t0 = np.searchsorted(tabs, a)
        ncols = np.searchsorted(tabs, b) - t0 + 1
        na = ncols < 5
        ok = ~na
        tabs_pad = np.concatenate([tabs, np.full(8, nb, dtype=np.int64)])

        def field(k):
            """(start, end) of column k of every line; only meaningful where ncols > k."""
            kk = np.asarray(k, dtype=np.int64) + np.zeros(n, dtype=np.int64)
            idx = np.minimum(t0 + kk, len(tabs_pad) - 1)
            s = np.where(kk == 0, a, tabs_pad[np.maximum(idx - 1, 0)] + 1)
            e = np.where(kk == ncols - 1, b, tabs_pad[idx])
            return s, e
        chrom_s, chrom_e = field(0)
        pos_s, pos_e = field(1)
        name_s, name_e = field(2)
        ref_s, ref_e = field(3)
        alt_s, alt_e = field(4)
        # ---- POS: decimal digits are parsed in bulk, anything else goes through int() like the reference (:98)
        pos = np.zeros(n, dtype=np.int64)
        plen = np.where(ok, pos_e - pos_s, 0)
        plain = ok & (plen > 0) & (plen <= 18)
        width = int(plen[plain].max()) if plain.any() else 0
        for k in range(width):
            sel = plain & (plen > k)
            d = buf[np.minimum(pos_s + k, nb - 1)].astype(np.int64) - 48
            plain &= ~(sel & ((d < 0) | (d > 9)))
            pos = np.where(sel, pos * 10 + d, pos)
        for i in np.flatnonzero(ok & ~plain):
            pos[i] = int(bytes(buf[pos_s[i]:pos_e[i]]).decode("utf-8", "replace"))     # ValueError as upstream
        # ---- REF '-' means empty (:100-102)
        dash = ok & (ref_e - ref_s == 1) & (buf[np.minimum(ref_s, nb - 1)] == 45)
        ref_e = np.where(dash, ref_s, ref_e)
        # ---- strand column (:105-111)
        minus = np.zeros(n, dtype=bool)
        if strand_index is not None:
            col = np.where(strand_index >= 0, strand_index, ncols + strand_index)
            bad = ok & ((col < 0) | (col >= ncols))
            if bad.any():
                raise IndexError("list index out of range")
            st_s, st_e = field(np.where(ok, col, 0))
            one = ok & (st_e - st_s == 1)
            ch = buf[np.minimum(st_s, nb - 1)]
            if require_strand:
                na = na | (one & (ch == 46))
            minus = one & (ch == 45)
        # ---- chromosome names: normalise the distinct ones (:84-96)
        ok = ok & ~na
        clen = np.where(ok, chrom_e - chrom_s, 0)
        width = int(clen.max()) if n else 0
        padded = np.zeros((n, max(width, 1)), dtype=np.uint8)
        dst0 = np.arange(n, dtype=np.int64) * max(width, 1)
        _ragged_copy(padded.reshape(-1), dst0, buf, chrom_s, clen)
        uniq, inv = np.unique(padded.view("S%d" % max(width, 1)).reshape(-1), return_inverse=True)
        names = [_normalise_chrom(u.decode("utf-8", "replace"), check_chr, genome_chrs) for u in uniq]
        canon = sorted(set(names))
        remap = np.array([canon.index(x) for x in names], dtype=np.int32)
        chrom_id = remap[inv.reshape(-1)] if n else np.zeros(0, np.int32)
        # ---- the window must lie inside the chromosome and off the blacklist (:113-121)
        if reference_sequence and seq_context:
            lhs, rhs = (seq_context, seq_context) if isinstance(seq_context, int) else seq_context
            half = (ref_e - ref_s) // 2
            start, end = pos + half - lhs, pos + half + rhs
            len_chrs = getattr(reference_sequence, "len_chrs", None)
            if isinstance(len_chrs, dict):
                # selene_b200.sequences.Genome: the checks of genome.py:51-92 on whole columns
                clens = np.array([len_chrs.get(c, -1) for c in canon], dtype=np.int64)
                cl = clens[chrom_id] if n else np.zeros(0, np.int64)
                inb = (cl >= 0) & (start < cl) & (start < end) & (end > 0) & (start >= 0) & (end <= cl)
                bl = getattr(reference_sequence, "_blacklist", None)
                if bl is not None:
                    for i in np.flatnonzero(ok & inb):
                        if bl.overlaps(canon[chrom_id[i]], int(start[i]), int(end[i])):
                            inb[i] = False
            else:
                # any other object with the reference's Sequence interface: ask it row by row
                inb = np.zeros(n, dtype=bool)
                for i in np.flatnonzero(ok):
                    inb[i] = bool(reference_sequence.coords_in_bounds(canon[chrom_id[i]], int(start[i]), int(end[i])))
            na = na | (ok & ~inb)
            ok = ok & inb
        if reference_sequence and seq_context and output_NAs_to_file:
            with open(output_NAs_to_file, 'w') as fh:
                ends = np.minimum(line_e + 1, nb)
                for i in np.flatnonzero(na):
                    text = bytes(buf[line_s[i]:ends[i]]).decode("utf-8", "replace")
                    fh.write(text.replace("\r\n", "\n").replace("\r", "\n"))
        # ---- ALT: '.' and ',' both separate alternatives (:122-124); one output row per alternative
        keep = np.flatnonzero(ok)
        delim = np.flatnonzero((buf == 44) | (buf == 46))
        d_lo = np.searchsorted(delim, alt_s[keep])
        d_hi = np.searchsorted(delim, alt_e[keep])
        n_alt = d_hi - d_lo + 1
        row = np.repeat(np.arange(len(keep), dtype=np.int64), n_alt)
        j = np.arange(len(row), dtype=np.int64) - np.repeat(np.cumsum(n_alt) - n_alt, n_alt)
        delim_pad = np.concatenate([delim, [nb]])
        src = keep[row]
        piece_s = np.where(j == 0, alt_s[src], delim_pad[np.maximum(d_lo[row] + j - 1, 0)] + 1)
        piece_e = np.where(j == n_alt[row] - 1, alt_e[src], delim_pad[np.minimum(d_lo[row] + j, len(delim_pad) - 1)])
        return cls(buf, canon, chrom_id[src], pos[src], name_s[src], name_e[src], ref_s[src], ref_e[src],
                   piece_s, piece_e, minus[src])


def read_vcf_file(input_path, strand_index=None, require_strand=False, output_NAs_to_file=None,
                  seq_context=None, reference_sequence=None):
    """The reference's entry point: list of (chrom, pos, id, ref, alt, strand) tuples."""
    return VariantTable.from_vcf(input_path, strand_index=strand_index, require_strand=require_strand,
                                 output_NAs_to_file=output_NAs_to_file, seq_context=seq_context,
                                 reference_sequence=reference_sequence).to_tuples()


def _get_ref_idxs(seq_len, ref_len):
    """Rows [start, end) of a window of ``seq_len`` that hold a reference allele of ``ref_len`` centred on the
    window's middle position (the lower middle for even lengths) — :137-143."""
    start = (seq_len - 1) // 2 - ref_len // 2
    return start, start + ref_len


def _eq_enc(a, b):
    """Equality of one-hot encodings expressed on strings: characters are equal as encodings iff
    they map to the same base, any two non-ACGT characters being the same 0.25 row."""
    if len(a) != len(b):
        return False
    norm = lambda ch: ch.upper() if ch in "ACGTacgt" else "N"
    return all(norm(x) == norm(y) for x, y in zip(a, b))


def _revcomp_enc_string(s):
    """get_reverse_complement_encoding (_common.py:37-62) on the string form of an encoding."""
    comp = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A', 'a': 't', 'c': 'g', 'g': 'c', 't': 'a'}
    return "".join(comp.get(ch, 'N') for ch in reversed(s))


def build_ref_alt_strings(variant, reference_sequence, sequence_length, start_radius, end_radius):
    """Ref and alt model-input sequences of one variant, any kind, as strings — the string-level
    restatement of model_predict.py:1054-1110 with ``_process_alt`` (:146-223),
    ``_handle_standard_ref`` (:226-245) and ``_handle_long_ref`` (:248-266).

    Returns (ref_string, alt_string, ref_match, contains_unk)."""
    chrom, pos, name, ref, alt, strand = variant
    center = pos + len(ref) // 2
    start = center - start_radius
    end = center + end_radius
    wt = reference_sequence.get_sequence_from_coords(chrom, start, end)
    contains_unk = reference_sequence.UNK_BASE in wt
    L = len(wt)
    # ---- _process_alt
    a = "" if alt in ('*', '-') else alt
    ref_len, alt_len = len(ref), len(a)
    if alt_len > L:
        s0 = int((alt_len - L) // 2)
        alt_seq = a[s0:s0 + L].upper()
    elif ref_len == alt_len:
        s, e = _get_ref_idxs(L, ref_len)
        alt_seq = wt[:s] + a + wt[e:]
    elif alt_len > ref_len:
        s, e = _get_ref_idxs(L, ref_len)
        seq = wt[:s] + a + wt[e:]
        trunc_s = (len(seq) - L) // 2
        alt_seq = seq[trunc_s:trunc_s + L]
    else:
        lhs = reference_sequence.get_sequence_from_coords(
            chrom, start - ref_len // 2 + alt_len // 2, pos + 1, pad=True)
        rhs = reference_sequence.get_sequence_from_coords(
            chrom, pos + 1 + ref_len, end + math.ceil(ref_len / 2.) - math.ceil(alt_len / 2.), pad=True)
        alt_seq = lhs + a + rhs
    # ---- reference handling
    match = True
    ref_seq = wt
    if ref_len and ref_len < sequence_length:
        s, _ = _get_ref_idxs(sequence_length, ref_len)
        match = _eq_enc(wt[s:s + ref_len], ref)
        if not match:
            ref_seq = wt[:s] + ref + wt[s + ref_len:]
    elif ref_len >= sequence_length:
        rs = ref_len // 2 - start_radius - 1
        re_ = ref_len // 2 + end_radius - 1
        cut = ref[rs:re_]
        match = _eq_enc(wt, cut)
        if not match:
            ref_seq = cut
    if strand == '-':
        ref_seq, alt_seq = _revcomp_enc_string(ref_seq), _revcomp_enc_string(alt_seq)
    return ref_seq, alt_seq, match, contains_unk
