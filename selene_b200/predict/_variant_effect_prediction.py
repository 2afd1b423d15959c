"""Variant ingest for variant effect prediction (reference:
selene_sdk/predict/_variant_effect_prediction.py).  SNVs — the data-parallel bulk — are turned into
(window start, ref code, alt code) arrays for ``sb_snv_pairs``; other variants get their ref/alt
windows built string-wise on the host exactly as the reference does and are scored from code rows.
"""
import math

VCF_REQUIRED_COLS = ["#CHROM", "POS", "ID", "REF", "ALT"]


def read_vcf_file(input_path, strand_index=None, require_strand=False, output_NAs_to_file=None,
                  seq_context=None, reference_sequence=None):
    """_variant_effect_prediction.py:13-134: list of (chrom, pos, id, ref, alt, strand)."""
    variants = []
    na_rows = []
    check_chr = True
    for chrom in reference_sequence.get_chrs():
        if not chrom.startswith("chr"):
            check_chr = False
            break
    with open(input_path, 'r') as file_handle:
        lines = file_handle.readlines()
        index = 0
        for index, line in enumerate(lines):
            if '#' not in line:
                break
            if "#CHROM" in line:
                cols = line.strip().split('\t')
                if cols[:5] != VCF_REQUIRED_COLS:
                    raise ValueError(
                        "First 5 columns in file {0} were {1}. "
                        "Expected columns: {2}".format(input_path, cols[:5], VCF_REQUIRED_COLS))
                index += 1
                break
        for line in lines[index:]:
            cols = line.strip().split('\t')
            if len(cols) < 5:
                na_rows.append(line)
                continue
            chrom = str(cols[0])
            if 'CHR' == chrom[:3]:
                chrom = chrom.replace('CHR', 'chr')
            elif "chr" not in chrom and check_chr is True:
                chrom = "chr" + chrom
            if chrom == "chrMT" and chrom not in reference_sequence.get_chrs():
                chrom = "chrM"
            elif chrom == "MT" and chrom not in reference_sequence.get_chrs():
                chrom = "M"
            pos = int(cols[1])
            name = cols[2]
            ref = cols[3]
            if ref == '-':
                ref = ""
            alt = cols[4]
            strand = '+'
            if strand_index is not None:
                if require_strand and cols[strand_index] == '.':
                    na_rows.append(line)
                    continue
                elif cols[strand_index] == '-':
                    strand = '-'
            if reference_sequence and seq_context:
                if isinstance(seq_context, int):
                    seq_context = (seq_context, seq_context)
                lhs_radius, rhs_radius = seq_context
                start = pos + len(ref) // 2 - lhs_radius
                end = pos + len(ref) // 2 + rhs_radius
                if not reference_sequence.coords_in_bounds(chrom, start, end):
                    na_rows.append(line)
                    continue
            alt = alt.replace('.', ',')  # consider '.' a valid delimiter
            for a in alt.split(','):
                variants.append((chrom, pos, name, ref, a, strand))
    if reference_sequence and seq_context and output_NAs_to_file:
        with open(output_NAs_to_file, 'w') as file_handle:
            for na_row in na_rows:
                file_handle.write(na_row)
    return variants


def _get_ref_idxs(seq_len, ref_len):
    """_variant_effect_prediction.py:137-143."""
    mid = seq_len // 2
    if seq_len % 2 == 0:
        mid -= 1
    start_pos = mid - ref_len // 2
    end_pos = start_pos + ref_len
    return (start_pos, end_pos)


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
