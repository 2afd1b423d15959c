"""CPU ORACLE — test infrastructure only, never shipped or measured as the product.

A plain numpy (and, for the one floating-point stage, torch-CPU fp32) restatement of the reference
algorithms on the sequence -> prediction hot path of FunctionLab/selene (selene-sdk 0.6.0).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; nothing under ``selene_b200/`` does.

Parity status: PINNED.  Every function below is checked in ``tests/test_oracle.py`` against
  * the reference's own known-answer tests (vectors lifted from sequences/tests/test_genome.py:40-106,
    predict/tests/test__common.py:22-48, predict/tests/test_model_predict.py:12-60,
    targets/tests/test_genomic_features.py:24-42,151-187), and
  * golden outputs produced by importing the unmodified reference in the authoring container
    (tests/golden/make_golden.py -> tests/golden/*.npz), and
  * when built, the reference's own two Cython modules compiled from where they lie
    (oracle/build_ref.py -> oracle/_ref/).

Each function cites the reference file:line it restates (paths relative to the reference root).
"""
import itertools
import math

import numpy as np

BASES_ARR = ['A', 'C', 'G', 'T']            # selene_sdk/sequences/genome.py:213
UNK_BASE = 'N'                              # selene_sdk/sequences/genome.py:248
COMPLEMENTARY_BASE_DICT = {                 # selene_sdk/sequences/genome.py:237-240
    'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A', 'N': 'N',
    'a': 'T', 'c': 'G', 'g': 'C', 't': 'A', 'n': 'N'}


def base_to_index(bases_arr=BASES_ARR):
    """selene_sdk/sequences/genome.py:279-286 (update_bases_order)."""
    d = {b: i for i, b in enumerate(bases_arr)}
    d.update({b.lower(): i for i, b in enumerate(bases_arr)})
    return d


# ------------------------------------------------------------------------------------------------
# stage 1a: one-hot encoding
# ------------------------------------------------------------------------------------------------
def sequence_to_encoding(sequence, bases_arr=BASES_ARR):
    """selene_sdk/sequences/_sequence.pyx:11-25 via sequences/sequence.py:14-41.

    encoding[i, base_to_index[c]] = 1; any character not in the dict -> whole row = float32(1/N).
    """
    b2i = base_to_index(bases_arr)
    n = len(bases_arr)
    enc = np.zeros((len(sequence), n), dtype=np.float32)
    fill = np.divide(1, n, dtype=np.float32)
    for i, ch in enumerate(sequence):
        if ch in b2i:
            enc[i, b2i[ch]] = 1
        else:
            enc[i, :] = fill
    return enc


def encoding_to_sequence(encoding, bases_arr=BASES_ARR, unk_base=UNK_BASE):
    """selene_sdk/sequences/sequence.py:44-85: per row, the first column equal to 1 names the base;
    a row without a 1 (e.g. 0.25 x 4) is the unknown base."""
    seq = []
    for row in encoding:
        pos = -1
        for index, val in enumerate(row):
            if val == 1:
                pos = index
                break
        seq.append(unk_base if pos == -1 else bases_arr[pos])
    return "".join(seq)


def check_coords(len_chrs, chrom, start, end, pad=False):
    """selene_sdk/sequences/genome.py:51-92 (_check_coords), no blacklist."""
    return (chrom in len_chrs and start < len_chrs[chrom] and start < end and end > 0 and
            (start >= 0 if not pad else True) and (end <= len_chrs[chrom] if not pad else True))


def revcomp_string(s):
    """pyfaidx ``.reverse.complement.seq`` as used by genome.py:348-352: case-preserving complement
    (A<->T, C<->G, a<->t, c<->g), every other character (N, n, ...) unchanged, then reversed."""
    comp = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A', 'a': 't', 'c': 'g', 'g': 'c', 't': 'a'}
    return "".join(comp.get(ch, ch) for ch in reversed(s))


def get_sequence_from_coords(genome, chrom, start, end, strand='+', pad=False):
    """selene_sdk/sequences/genome.py:96-166 with ``genome`` a dict chrom -> str.

    Invalid coordinates -> "".  pad=True N-fills out-of-range parts; note the padding is added
    AFTER the strand is applied to the in-bounds part (genome.py:164-166)."""
    len_chrs = {k: len(v) for k, v in genome.items()}
    if not check_coords(len_chrs, chrom, start, end, pad=pad):
        return ""
    if strand not in ('+', '-', '.'):
        raise ValueError("Strand must be one of '+', '-', or '.'. Input was {0}".format(strand))
    end_pad = start_pad = 0
    if end > len_chrs[chrom]:
        end_pad = end - len_chrs[chrom]
        end = len_chrs[chrom]
    if start < 0:
        start_pad = -start
        start = 0
    inner = genome[chrom][start:end]
    if strand == '-':
        inner = revcomp_string(inner)
    return UNK_BASE * start_pad + inner + UNK_BASE * end_pad


def get_encoding_from_coords_check_unk(genome, chrom, start, end, strand='+', pad=False,
                                       bases_arr=BASES_ARR):
    """selene_sdk/sequences/genome.py:488-542: (encoding, 'N' in sequence) — uppercase N only."""
    s = get_sequence_from_coords(genome, chrom, start, end, strand=strand, pad=pad)
    return sequence_to_encoding(s, bases_arr), (UNK_BASE in s)


# ------------------------------------------------------------------------------------------------
# stage 1b: interval-overlap labels
# ------------------------------------------------------------------------------------------------
def query_rows(rows_by_chrom, chrom, start, end):
    """tabix ``query(chrom, start, end)`` on a BED file as genomic_features.py:341-348 sees it:
    rows whose [row_start, row_end) overlaps [start, end): row_start < end and row_end > start.
    Returns None for an unknown contig (the TabixError branch, :347-348)."""
    if chrom not in rows_by_chrom:
        return None
    return [r for r in rows_by_chrom[chrom] if int(r[1]) < end and int(r[2]) > start]


def fast_get_feature_data(start, end, thresholds, feature_index_dict, rows):
    """selene_sdk/targets/_genomic_features.pyx:12-41."""
    n_features = len(feature_index_dict)
    query_length = end - start
    encoding = np.zeros((query_length, n_features), dtype=np.int64)
    if rows is None:
        return np.zeros((n_features,))
    for row in rows:
        feature_start = int(row[1])
        feature_end = int(row[2])
        index_start = max(0, feature_start - start)
        index_end = min(feature_end - start, query_length)
        index_feat = feature_index_dict[row[3]]
        if index_start == index_end:
            index_end += 1
        encoding[index_start:index_end, index_feat] = 1
    thresholds = (np.asarray(thresholds, dtype=np.float32) * query_length - 1).clip(min=0)
    return (np.sum(encoding, axis=0) > thresholds.astype(np.int64)).astype(np.int64)


def integer_thresholds(thresholds, query_length):
    """The integer the reference compares coverage with (_genomic_features.pyx:39-40): float32
    arithmetic, clip at 0, truncate."""
    t = (np.asarray(thresholds, dtype=np.float32) * query_length - 1).clip(min=0)
    return t.astype(np.int64)


def feature_data_no_threshold(n_features, feature_index_dict, rows):
    """selene_sdk/targets/genomic_features.py:406-415: any returned row marks its feature (last
    column) with 1.0; float64 output."""
    features = np.zeros(n_features)
    if not rows:
        return features
    for r in rows:
        features[feature_index_dict[r[-1]]] = 1
    return features


# ------------------------------------------------------------------------------------------------
# stage 2a: mutant generation
# ------------------------------------------------------------------------------------------------
def in_silico_mutagenesis_sequences(sequence, mutate_n_bases=1, bases_arr=BASES_ARR,
                                    start_position=0, end_position=None):
    """selene_sdk/predict/_in_silico_mutagenesis.py:8-107 (argument checks :64-89 included)."""
    if end_position is None:
        end_position = len(sequence)
    if start_position >= end_position:
        raise ValueError("Starting positions must be less than the ending positions.")
    if start_position < 0:
        raise ValueError("Negative starting positions are not supported.")
    if end_position < 0:
        raise ValueError("Negative ending positions are not supported.")
    if start_position >= len(sequence):
        raise ValueError("Starting positions must be less than the sequence length.")
    if end_position > len(sequence):
        raise ValueError("Ending positions must be less than or equal to the sequence length.")
    if (end_position - start_position) < mutate_n_bases:
        raise ValueError("Fewer bases exist in the substring than need to be mutated.")
    sequence_alts = [[b for b in bases_arr if b != ref] for ref in sequence]
    out = []
    for indices in itertools.combinations(range(start_position, end_position), mutate_n_bases):
        for mutations in itertools.product(*[sequence_alts[i] for i in indices]):
            out.append(list(zip(indices, mutations)))
    return out


def mutate_sequence(encoding, mutation_information, bases_arr=BASES_ARR):
    """selene_sdk/predict/_in_silico_mutagenesis.py:110-143."""
    b2i = base_to_index(bases_arr)
    mutated = np.copy(encoding)
    for position, alt in mutation_information:
        mutated[position, :] = 0
        mutated[position, b2i[alt]] = 1
    return mutated


def get_ref_idxs(seq_len, ref_len):
    """selene_sdk/predict/_variant_effect_prediction.py:137-143."""
    mid = seq_len // 2
    if seq_len % 2 == 0:
        mid -= 1
    start_pos = mid - ref_len // 2
    return start_pos, start_pos + ref_len


def radii(sequence_length):
    """selene_sdk/predict/model_predict.py:168-171."""
    start_radius = sequence_length // 2
    end_radius = start_radius
    if sequence_length % 2 != 0:
        start_radius += 1
    return start_radius, end_radius


def get_reverse_complement_encoding(allele_encoding, bases_arr=BASES_ARR,
                                    complementary_base_dict=COMPLEMENTARY_BASE_DICT):
    """selene_sdk/predict/_common.py:37-62."""
    base_ixs = {b: i for i, b in enumerate(bases_arr)}
    complement_indices = [base_ixs[complementary_base_dict[b]] for b in bases_arr]
    return allele_encoding[:, complement_indices][::-1, :]


def snv_ref_alt_encodings(genome, chrom, pos, ref, alt, strand, sequence_length, bases_arr=BASES_ARR):
    """The per-variant body of selene_sdk/predict/model_predict.py:1054-1110 restricted to
    substitutions with len(ref) == len(alt) >= 1 (``_process_alt`` substitution branch,
    _variant_effect_prediction.py:195-200; ``_handle_standard_ref``, :226-245).

    Returns (ref_encoding, alt_encoding, ref_match, contains_unk)."""
    start_radius, end_radius = radii(sequence_length)
    center = pos + len(ref) // 2
    start, end = center - start_radius, center + end_radius
    wt, contains_unk = get_encoding_from_coords_check_unk(genome, chrom, start, end, bases_arr=bases_arr)
    ref_encoding = sequence_to_encoding(ref, bases_arr)
    alt_encoding = sequence_to_encoding(alt, bases_arr)
    s, e = get_ref_idxs(len(wt), len(ref))
    alt_seq = np.vstack([wt[:s, :], alt_encoding, wt[e:, :]])
    ref_seq = wt  # the reference mutates the array it got back in place
    match = np.array_equal(ref_seq[s:s + len(ref), :], ref_encoding)
    if not match:
        ref_seq[s:s + len(ref), :] = ref_encoding
    if strand == '-':
        ref_seq = get_reverse_complement_encoding(ref_seq, bases_arr)
        alt_seq = get_reverse_complement_encoding(alt_seq, bases_arr)
    return ref_seq, alt_seq, bool(match), bool(contains_unk)


# ------------------------------------------------------------------------------------------------
# stage 2b: DeepSEA-family forward (torch CPU fp32, the same ops the reference dispatches)
# ------------------------------------------------------------------------------------------------
def stack_forward(x_bl4, layers):
    """``predict()`` (selene_sdk/predict/_common.py:65-97) + a conv stack such as
    models/deepsea.py:21-58.  ``x_bl4``: (B, L, 4) array; ``layers``: list of dicts
    {type:'conv'|'fc', weight, bias, relu, pool} in torch layouts.  Returns float32 (B, F) sigmoid
    outputs.  Dropout is identity in eval mode; flatten is channel-major (deepsea.py:56)."""
    import torch
    import torch.nn.functional as F
    with torch.no_grad():
        x = torch.tensor(np.asarray(x_bl4, dtype=np.float32)).transpose(1, 2)
        flat = False
        for lay in layers:
            w = torch.tensor(np.asarray(lay["weight"], dtype=np.float32))
            b = torch.tensor(np.asarray(lay["bias"], dtype=np.float32))
            if lay["type"] == "conv":
                x = F.conv1d(x, w, b)
                if lay.get("relu"):
                    x = F.relu(x)
                if lay.get("pool", 1) > 1:
                    x = F.max_pool1d(x, lay["pool"], lay["pool"])
            else:
                if not flat:
                    x = x.reshape(x.size(0), -1)
                    flat = True
                x = F.linear(x, w, b)
                if lay.get("relu"):
                    x = F.relu(x)
        return torch.sigmoid(x).numpy()


def deepsea_layers(sequence_length=1000, n_features=919, seed=0, scale=1.0):
    """Random-init DeepSEA (models/deepsea.py:21-50) weights as a layer list; torch default init
    under ``torch.manual_seed(seed)`` in module construction order.  ``scale`` multiplies the conv
    weights (default init gives ~0.5 outputs; scaling exercises saturation)."""
    import torch
    import torch.nn as nn
    torch.manual_seed(seed)
    c1, c2, c3 = nn.Conv1d(4, 320, 8), nn.Conv1d(320, 480, 8), nn.Conv1d(480, 960, 8)
    n_channels = int(np.floor((np.floor((sequence_length - 7) / 4.) - 7) / 4.) - 7)
    f1, f2 = nn.Linear(960 * n_channels, n_features), nn.Linear(n_features, n_features)

    def g(m, s=1.0):
        return (m.weight.detach().numpy().copy() * s).astype(np.float32), m.bias.detach().numpy().copy()
    layers = []
    for m, pool in ((c1, 4), (c2, 4), (c3, 1)):
        w, b = g(m, scale)
        layers.append(dict(type="conv", weight=w, bias=b, relu=1, pool=pool))
    w, b = g(f1)
    layers.append(dict(type="fc", weight=w, bias=b, relu=1, pool=1))
    w, b = g(f2)
    layers.append(dict(type="fc", weight=w, bias=b, relu=0, pool=1))
    return layers


# ------------------------------------------------------------------------------------------------
# stage 3: scores
# ------------------------------------------------------------------------------------------------
def abs_diff_score(batch_predictions, baseline_predictions):
    """predict_handlers/absolute_diff_score_handler.py:115."""
    return np.abs(baseline_predictions - batch_predictions)


def diff_score(batch_predictions, baseline_predictions):
    """predict_handlers/diff_score_handler.py:114."""
    return batch_predictions - baseline_predictions


def logit_score(batch_predictions, baseline_predictions):
    """predict_handlers/logit_score_handler.py:118-124 — clamps BOTH inputs in place, then
    scipy.special.logit(alt) - logit(ref) in float32."""
    from scipy.special import logit
    baseline_predictions[baseline_predictions == 0] = 1e-24
    baseline_predictions[baseline_predictions >= 1] = 0.999999
    batch_predictions[batch_predictions == 0] = 1e-24
    batch_predictions[batch_predictions >= 1] = 0.999999
    return logit(batch_predictions) - logit(baseline_predictions)


# ------------------------------------------------------------------------------------------------
# whole-path helpers used by bench.py's CPU legs and by smoke()
# ------------------------------------------------------------------------------------------------
def ism_scores(sequence, layers, batch_size=64, bases_arr=BASES_ARR):
    """In-silico mutagenesis of one sequence as selene_sdk/predict/model_predict.py:599-660,749-781
    (single-base), returning (mutations, abs_diffs, logits, base_prediction)."""
    sequence = sequence.upper()
    muts = in_silico_mutagenesis_sequences(sequence, 1, bases_arr)
    enc = sequence_to_encoding(sequence, bases_arr)
    base = stack_forward(enc.reshape(1, *enc.shape), layers)
    absd, logits = [], []
    for i in range(0, len(muts), batch_size):
        chunk = muts[i:i + batch_size]
        batch = np.zeros((len(chunk), enc.shape[0], enc.shape[1]))
        for j, m in enumerate(chunk):
            batch[j] = mutate_sequence(enc, m, bases_arr)
        out = stack_forward(batch, layers)
        absd.append(abs_diff_score(out, base))
        logits.append(logit_score(out, base))
    return muts, np.concatenate(absd), np.concatenate(logits), base


def vep_scores(genome, variants, layers, sequence_length, batch_size=64, bases_arr=BASES_ARR):
    """SNV variant effect prediction as selene_sdk/predict/model_predict.py:1054-1138 with
    abs_diffs + logits.  ``variants``: list of (chrom, pos, ref, alt, strand).
    Returns (abs_diffs, logits, ref_match, contains_unk)."""
    absd, logits, matches, unks = [], [], [], []
    refs, alts = [], []

    def flush():
        r = stack_forward(np.array(refs), layers)
        a = stack_forward(np.array(alts), layers)
        absd.append(abs_diff_score(a, r))
        logits.append(logit_score(a, r))
        del refs[:], alts[:]
    for chrom, pos, ref, alt, strand in variants:
        r, a, m, u = snv_ref_alt_encodings(genome, chrom, pos, ref, alt, strand, sequence_length, bases_arr)
        refs.append(r)
        alts.append(a)
        matches.append(m)
        unks.append(u)
        if len(refs) >= batch_size:
            flush()
    if refs:
        flush()
    return np.concatenate(absd), np.concatenate(logits), np.array(matches), np.array(unks)


# ------------------------------------------------------------------------------------------------
# stage 1 caller: RandomPositionsSampler (training config)
# ------------------------------------------------------------------------------------------------
def get_indices_and_probabilities(interval_lengths, indices):
    """selene_sdk/utils/utils.py:34-67."""
    select = np.array(interval_lengths)[indices]
    weights = select / float(np.sum(select))
    keep = [indices[i] for i, w in enumerate(weights) if w > 1e-10]
    if len(keep) == len(indices):
        return indices, weights.tolist()
    return get_indices_and_probabilities(interval_lengths, keep)


class RandomPositionsSamplerOracle(object):
    """selene_sdk/samplers/random_positions_sampler.py:120-369 + online_sampler.py:121-216,
    chromosome-holdout mode, on a dict genome and BED rows.  RNG streams exactly as the reference:
    ``np.random.seed(seed)``, ``random.seed(seed + 1)`` (online_sampler.py:136-138); one
    ``np.random.choice(indices, 200000, p=weights)`` cache per mode at first use for ALL modes
    (:175-176,306-314); per draw ``np.random.randint(cstart, cend)`` then, after the label query,
    ``random.randint(0, 1)`` for the strand (:277); rejected draws consume both."""
    STRAND_SIDES = ('+', '-')

    def __init__(self, genome, rows_by_chrom, features, seed=436, validation_holdout=('chr6', 'chr7'),
                 test_holdout=('chr8', 'chr9'), exclude_chrs=('_',), sequence_length=1000,
                 center_bin_to_predict=200, feature_thresholds=0.5, mode="train"):
        import random
        self._random = random
        np.random.seed(seed)
        random.seed(seed + 1)
        self.genome, self.rows = genome, rows_by_chrom
        self.features = list(features)
        self.fid = {f: i for i, f in enumerate(self.features)}
        self.thr = np.full(len(self.features), feature_thresholds, dtype=np.float32)
        self.sequence_length = sequence_length
        self.modes = ["train", "validate"] + (["test"] if test_holdout else [])
        self.mode = mode
        wr = int(sequence_length / 2)
        self._start_window_radius, self._end_window_radius = wr, wr + (1 if sequence_length % 2 else 0)
        br = int(center_bin_to_predict / 2)
        self._start_radius, self._end_radius = br, br + (1 if center_bin_to_predict % 2 else 0)
        by_mode = {m: [] for m in self.modes}
        self.sample_from_intervals, self.interval_lengths = [], []
        index = 0
        for chrom in sorted(genome.keys()):
            n = len(genome[chrom])
            if any(e in chrom for e in exclude_chrs):
                continue
            if chrom in validation_holdout:
                by_mode["validate"].append(index)
            elif test_holdout and chrom in test_holdout:
                by_mode["test"].append(index)
            else:
                by_mode["train"].append(index)
            self.sample_from_intervals.append((chrom, sequence_length, n - sequence_length))
            self.interval_lengths.append(n - 2 * sequence_length)
            index += 1
        self._from_mode = {m: get_indices_and_probabilities(self.interval_lengths, by_mode[m]) for m in self.modes}
        self._cache = {}
        for m in self.modes:
            self._update_randcache(m)

    def _update_randcache(self, mode):
        ind, w = self._from_mode[mode]
        self._cache[mode] = {"idx": np.random.choice(ind, size=200000, replace=True, p=w), "next": 0}

    def _retrieve(self, chrom, position):
        bs, be = position - self._start_radius, position + self._end_radius
        targets = fast_get_feature_data(bs, be, self.thr, self.fid, query_rows(self.rows, chrom, bs, be))
        ws, we = position - self._start_window_radius, position + self._end_window_radius
        strand = self.STRAND_SIDES[self._random.randint(0, 1)]
        seq = sequence_to_encoding(get_sequence_from_coords(self.genome, chrom, ws, we, strand))
        if seq.shape[0] == 0 or np.mean(seq == 0.25) > 0.30:
            return None
        return seq, targets

    def sample(self, batch_size=1, mode=None):
        mode = mode or self.mode
        sequences = np.zeros((batch_size, self.sequence_length, 4))
        targets = np.zeros((batch_size, len(self.features)))
        n = 0
        while n < batch_size:
            c = self._cache[mode]
            if c["next"] == len(c["idx"]):
                self._update_randcache(self.mode)   # the reference refreshes self.mode's cache (:352)
                c = self._cache[mode]
                c["next"] = 0
            k = c["idx"][c["next"]]
            c["next"] += 1
            chrom, cstart, cend = self.sample_from_intervals[k]
            position = np.random.randint(cstart, cend)
            out = self._retrieve(chrom, position)
            if not out:
                continue
            sequences[n], targets[n] = out
            n += 1
        return sequences, targets


# ------------------------------------------------------------------------------------------------
# packed training samples
# ------------------------------------------------------------------------------------------------
def pack_samples(sequences, targets):
    """scripts/sampler_write_to_h5.py:63-65 (--compress): bits along the position / target axis, MSB first."""
    return np.packbits(np.asarray(sequences) > 0, axis=1), np.packbits(np.asarray(targets) > 0, axis=1)


def unpackbits_sequence(sequence, s_len):
    """selene_sdk/samplers/dataloader.py:129-138."""
    sequence = np.unpackbits(sequence.astype(np.uint8), axis=-2)
    nulls = np.sum(sequence, axis=-1) == sequence.shape[-1]
    sequence = sequence.astype(float)
    sequence[nulls, :] = 1.0 / sequence.shape[-1]
    return sequence[:, :s_len, :] if sequence.ndim == 3 else sequence[:s_len, :]


def unpackbits_targets(targets, t_len):
    """selene_sdk/samplers/dataloader.py:140-146 (its 1-D branch refers to an undefined ``self``; only the 2-D form is
    reachable from the DataLoader path and the 1-D case is the obvious slice)."""
    targets = np.unpackbits(targets, axis=-1).astype(float)
    return targets[:, :t_len] if targets.ndim == 2 else targets[:t_len]
