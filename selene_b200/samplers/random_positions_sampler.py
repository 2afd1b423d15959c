"""``RandomPositionsSampler`` — drop-in for ``selene_sdk.samplers.RandomPositionsSampler`` (reference:
samplers/random_positions_sampler.py:23-369 on top of samplers/online_sampler.py:22-464) whose
mini-batches are produced by two device calls instead of a per-sample Python loop.

What stays on the host, unchanged: the RNG streams and their call order (``np.random.seed(seed)``,
``random.seed(seed+1)``; one 200 000-entry ``np.random.choice`` cache per mode; per draw
``np.random.randint`` for the position, then ``random.randint(0, 1)`` for the strand; rejected draws
consume both) and the rejection test (empty window, or more than 30 % unknown bases) — the latter is
answered from a host prefix sum of the unknown mask, so a whole batch of accepted (chrom, position,
strand) triples is known before the device is touched.  What moves to the GPU: the B window
encodings (``sb_encode_windows``) and the B x F labels (``sb_label_bins``).

``sample()`` returns the reference's float64 numpy arrays; ``sample_device()`` returns CUDA tensors
(float32 / bf16 sequences, uint8 labels) for a training loop that never leaves the device.
"""
import random
from collections import namedtuple
from functools import wraps

import numpy as np

from ..targets import GenomicFeatures

SampleIndices = namedtuple("SampleIndices", ["indices", "weights"])


def get_indices_and_probabilities(interval_lengths, indices):
    """Sampling weights of the intervals ``indices``, proportional to their lengths; intervals whose share is
    <= 1e-10 are dropped and the weights renormalised until none is (selene_sdk/utils/utils.py:34-67)."""
    lengths = np.array(interval_lengths)
    while True:
        sel = lengths[indices]
        weights = sel / float(np.sum(sel))
        keep = [idx for idx, w in zip(indices, weights) if w > 1e-10]
        if len(keep) == len(indices):
            return indices, weights.tolist()
        indices = keep


class RandomPositionsSampler(object):
    """Constructor arguments as random_positions_sampler.py:123-136.  ``reference_sequence`` is a
    :class:`selene_b200.sequences.Genome`; ``target_path`` a BED(.gz) file."""

    STRAND_SIDES = ('+', '-')
    BASE_MODES = ("train", "validate")

    def __init__(self, reference_sequence, target_path, features, seed=436,
                 validation_holdout=['chr6', 'chr7'], test_holdout=['chr8', 'chr9'], exclude_chrs=['_'],
                 sequence_length=1000, center_bin_to_predict=200, feature_thresholds=0.5, mode="train",
                 save_datasets=[], output_dir=None):
        self._features = features
        self.n_features = len(features)
        self.seed = seed
        np.random.seed(seed)           # the two RNG streams of online_sampler.py:136-138, in this order
        random.seed(seed + 1)
        self.sequence_length = sequence_length
        (self._start_window_radius, self._end_window_radius,
         self._start_radius, self._end_radius) = self._radii(sequence_length, center_bin_to_predict)
        self._holdout_type, self.validation_holdout, self.test_holdout = self._holdouts(validation_holdout,
                                                                                        test_holdout)
        self.modes = list(self.BASE_MODES) + (["test"] if self.test_holdout else [])
        if mode not in self.modes:
            raise ValueError("Mode must be one of {0}. Input was '{1}'.".format(self.modes, mode))
        self.mode = mode
        self.reference_sequence = reference_sequence
        self.target = (GenomicFeatures(target_path, features, feature_thresholds=feature_thresholds)
                       if isinstance(target_path, str) else target_path)
        self.exclude_chrs = exclude_chrs
        self._save_datasets = {m: [] for m in save_datasets}
        self._output_dir = output_dir
        self._sample_from_mode = dict.fromkeys(self.modes)
        self._randcache = {m: {"cache_indices": None, "sample_next": 0} for m in self.modes}
        self.sample_from_intervals = []
        self.interval_lengths = []
        self._initialized = False

    @staticmethod
    def _radii(sequence_length, center_bin):
        """(window start radius, window end radius, bin start radius, bin end radius) around a sampled position:
        the odd base goes to the end side; a ``[start, end]`` bin is converted to radii (online_sampler.py:188-216)."""
        if isinstance(center_bin, int) and (sequence_length + center_bin) % 2 != 0:
            raise ValueError(
                "Sequence length of {0} with a center bin length of {1} "
                "is invalid. These 2 inputs should both be odd or both be "
                "even.".format(sequence_length, center_bin))
        w_start = sequence_length // 2
        w_end = w_start + sequence_length % 2
        if isinstance(center_bin, int):
            return w_start, w_end, center_bin // 2, center_bin // 2 + center_bin % 2
        if not isinstance(center_bin, list) or len(center_bin) != 2:
            raise ValueError(
                "`center_bin_to_predict` needs to be either an int or a list of "
                "two ints, but was type '{0}'".format(type(center_bin)))
        bin_start, bin_end = center_bin
        return w_start, w_end, w_start - bin_start, w_end - (sequence_length - bin_end)

    @staticmethod
    def _holdouts(validation, test):
        """("chromosome" | "proportion", validation holdout, test holdout or None); lists of names hold out whole
        chromosomes, floats a share of them, and the two must agree in kind (online_sampler.py:147-186)."""
        kind = lambda h: "chromosome" if isinstance(h, list) else ("proportion" if isinstance(h, float) else None)
        norm = lambda h: [str(c) for c in h] if isinstance(h, list) else h
        if not test:
            if kind(validation) is None:
                raise ValueError(
                    "Validation holdout must be of type list (chromosomal "
                    "holdout) or float (proportion holdout) but was type "
                    "{0}.".format(type(validation)))
            return kind(validation), norm(validation), None
        if kind(validation) is None or kind(validation) != kind(test):
            raise ValueError(
                "Validation holdout and test holdout must have the "
                "same type (list or float) but validation was "
                "type {0} and test was type {1}".format(type(validation), type(test)))
        return kind(validation), norm(validation), norm(test)

    # -- lazy partitioning of the genome into the modes' sampling intervals (random_positions_sampler.py:166-262) --
    def init(func):
        @wraps(func)
        def dfunc(self, *args, **kwargs):
            if not self._initialized:
                self._partition_genome()
                for mode in self.modes:
                    self._update_randcache(mode=mode)
                self._initialized = True
            return func(self, *args, **kwargs)
        return dfunc

    def _skip(self, chrom):
        return any(excl in chrom for excl in self.exclude_chrs)

    def _partition_genome(self):
        """One sampling interval ``[L, len - L)`` per kept chromosome, assigned to a mode either by name or by a
        shuffled split (one ``np.random.shuffle`` call, as upstream); modes draw intervals with probability
        proportional to ``interval_lengths`` (which upstream fills with ``len - 2L`` for chromosome holdouts and
        with ``len`` for proportional ones)."""
        L = self.sequence_length
        by_name = self._holdout_type == "chromosome"
        kept = [(c, n) for c, n in self.reference_sequence.get_chr_lens() if not self._skip(c)]
        self.sample_from_intervals = [(c, L, n - L) for c, n in kept]
        self.interval_lengths = [n - 2 * L if by_name else n for c, n in kept]
        members = {m: [] for m in self.modes}
        if by_name:
            for i, (c, _n) in enumerate(kept):
                if c in self.validation_holdout:
                    members["validate"].append(i)
                elif self.test_holdout and c in self.test_holdout:
                    members["test"].append(i)
                else:
                    members["train"].append(i)
        else:
            order = list(range(len(kept)))
            np.random.shuffle(order)
            n_val = int(len(kept) * self.validation_holdout)
            n_test = int(len(kept) * self.test_holdout) if self.test_holdout else 0
            members["validate"] = order[:n_val]
            if self.test_holdout:
                members["test"] = order[n_val:n_val + n_test]
            members["train"] = order[n_val + n_test:]
        # upstream builds validate, then test, then train for proportional holdouts and iterates ``modes`` otherwise;
        # no RNG is involved, so the order is immaterial
        for m in self.modes:
            self._sample_from_mode[m] = SampleIndices(*get_indices_and_probabilities(self.interval_lengths, members[m]))

    def _update_randcache(self, mode=None):
        if not mode:
            mode = self.mode
        self._randcache[mode]["cache_indices"] = np.random.choice(
            self._sample_from_mode[mode].indices, size=200000, replace=True,
            p=self._sample_from_mode[mode].weights)
        self._randcache[mode]["sample_next"] = 0

    # -- the draw loop, host side only ---------------------------------------------------------------
    def _draw(self, batch_size, mode):
        """Accepted (chrom, position, strand) triples of one mini-batch, consuming the RNG streams in
        the reference's order (random_positions_sampler.py:348-368 with _retrieve, :264-304)."""
        G = self.reference_sequence
        chroms, positions, strands = [], [], []
        while len(chroms) < batch_size:
            cache = self._randcache[mode]
            sample_index = cache["sample_next"]
            if sample_index == len(cache["cache_indices"]):
                self._update_randcache()          # refreshes self.mode's cache, as the reference does
                cache = self._randcache[mode]
                sample_index = 0
            rand_interval_index = cache["cache_indices"][sample_index]
            cache["sample_next"] += 1
            chrom, cstart, cend = self.sample_from_intervals[rand_interval_index]
            position = np.random.randint(cstart, cend)
            window_start = position - self._start_window_radius
            window_end = position + self._end_window_radius
            if window_end - window_start < self.sequence_length:
                continue
            strand = self.STRAND_SIDES[random.randint(0, 1)]
            if not G.coords_in_bounds(chrom, window_start, window_end):
                continue                          # empty encoding -> sample again
            n_unk = G.count_unknown(chrom, window_start, window_end)
            if n_unk / float(window_end - window_start) > 0.30:
                continue                          # np.mean(seq == 0.25) > 0.30
            chroms.append(chrom)
            positions.append(position)
            strands.append(strand)
        return chroms, np.asarray(positions, dtype=np.int64), strands

    @init
    def sample_device(self, batch_size=1, mode=None, seq_dtype="float32", label_dtype="float32"):
        """One mini-batch as CUDA tensors: sequences (B, L, 4), targets (B, F)."""
        mode = mode if mode else self.mode
        chroms, positions, strands = self._draw(batch_size, mode)
        seqs, _ = self.reference_sequence.encode_windows(
            chroms, positions - self._start_window_radius, self.sequence_length, strands=strands, dtype=seq_dtype)
        targets = self.target.label_bins(chroms, positions - self._start_radius, positions + self._end_radius,
                                         out_dtype=label_dtype)
        if mode in self._save_datasets:
            t_host = targets.cpu().numpy()
            for c, p, s, t in zip(chroms, positions, strands, t_host):
                self._save_datasets[mode].append(
                    [c, int(p) - self._start_window_radius, int(p) + self._end_window_radius, s,
                     ';'.join(str(f) for f in np.nonzero(t)[0])])
        return seqs, targets

    @init
    def sample(self, batch_size=1, mode=None):
        """random_positions_sampler.py:316-369: float64 (B, L, 4) and float64 (B, F) numpy arrays."""
        seqs, targets = self.sample_device(batch_size, mode)
        return seqs.cpu().numpy().astype(np.float64), targets.cpu().numpy().astype(np.float64)

    # -- the rest of the Sampler / OnlineSampler surface ----------------------------------------------
    def set_mode(self, mode):
        if mode not in self.modes:
            raise ValueError("Tried to set mode to be '{0}' but the only valid modes are {1}".format(mode, self.modes))
        self.mode = mode

    def get_feature_from_index(self, index):
        return self.target.index_feature_dict[index]

    def get_sequence_from_encoding(self, encoding):
        return self.reference_sequence.encoding_to_sequence(encoding)

    def get_data_and_targets(self, batch_size, n_samples=None, mode=None):
        """online_sampler.py:297-354."""
        if mode is not None:
            self.set_mode(mode)
        else:
            mode = self.mode
        sequences_and_targets = []
        if n_samples is None and mode == "validate":
            n_samples = 32000
        elif n_samples is None and mode == "test":
            n_samples = 640000
        n_batches = int(n_samples / batch_size)
        for _ in range(n_batches):
            inputs, targets = self.sample(batch_size)
            sequences_and_targets.append((inputs, targets))
        targets_mat = np.vstack([t for (s, t) in sequences_and_targets])
        return sequences_and_targets, targets_mat

    def get_validation_set(self, batch_size, n_samples=None):
        return self.get_data_and_targets(batch_size, n_samples, mode="validate")

    def get_test_set(self, batch_size, n_samples=None):
        if "test" not in self.modes:
            raise ValueError("No test partition of the data was specified during initialization. "
                             "Cannot use method `get_test_set`.")
        return self.get_data_and_targets(batch_size, n_samples, mode="test")

    def save_dataset_to_file(self, mode, close_filehandle=False):
        """online_sampler.py:265-295: appends the sampled (chrom, start, end, strand, features) rows."""
        import os
        if mode not in self._save_datasets or not self._save_datasets[mode]:
            return
        path = os.path.join(self._output_dir or ".", "{0}_data.bed".format(mode))
        with open(path, "a") as fh:
            for cols in self._save_datasets[mode]:
                fh.write("{0}\n".format("\t".join(str(c) for c in cols)))
        self._save_datasets[mode] = []
