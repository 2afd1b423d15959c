"""Packed training samples on the device — SURVEY §8 f4 (reference: selene_sdk/samplers/dataloader.py:129-385,
scripts/sampler_write_to_h5.py:63-89).

The reference stores pre-sampled training data bit-packed (``np.packbits(sequences > 0, axis=1)``: one bit per
(position, channel), unknown bases as four set bits; ``np.packbits(targets > 0, axis=1)``) and unpacks one sample at
a time in DataLoader workers (``unpackbits_sequence`` / ``unpackbits_targets``).  Here the packed arrays are uploaded
once (0.5 B per base: 4 M samples of 1000 bp are 2 GB, resident in HBM) and a batch is ``sb_unpack_sequences`` +
``sb_unpack_targets`` on the gathered rows — straight to the (B, L, 4) / (B, F) tensors the training step reads.

``unpackbits_sequence`` / ``unpackbits_targets`` keep the reference's numpy signatures; ``H5Dataset`` /
``H5DataLoader`` keep its constructor arguments.  The file is read with ``h5py`` when that package is installed;
any mapping with the same keys (``sequences``, ``targets``, ``sequences_length``, ``targets_length`` — e.g. an
``np.load`` of an ``.npz``) is accepted in its place, which is also what the tests use (h5py is not in this image).
"""
import numpy as np

from .. import _lib


def _cuda(a, dtype=None):
    import torch
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def unpack_sequences_device(packed, s_len, seq_start=0, out_len=None, dtype="float32"):
    """packed: uint8 (n, ceil(L/8), 4) (numpy or CUDA tensor) -> CUDA (n, out_len, 4) fp32 / bf16."""
    import torch
    p = _cuda(packed, torch.uint8)
    n = int(p.shape[0])
    stored = int(p.shape[1]) * 8
    s_len = int(s_len)
    out_len = s_len - seq_start if out_len is None else int(out_len)
    if p.shape[-1] != 4:
        raise ValueError("packed sequences must have 4 channels in the last axis, got shape %s" % (tuple(p.shape),))
    if not (0 <= seq_start and seq_start + out_len <= min(s_len, stored)):
        raise ValueError("positions [%d, %d) are outside the %d stored positions" % (seq_start, seq_start + out_len, s_len))
    out = torch.empty((n, out_len, 4), dtype=torch.float32 if dtype == "float32" else torch.bfloat16, device=p.device)
    _lib.check(_lib.load().sb_unpack_sequences(_lib.ptr(p), n, stored, int(seq_start), out_len,
                                               _lib.SB_F32 if dtype == "float32" else _lib.SB_BF16, _lib.ptr(out),
                                               _lib.stream_ptr()), "sb_unpack_sequences")
    return out


def unpack_targets_device(packed, t_len, dtype="float32"):
    """packed: uint8 (n, ceil(F/8)) -> CUDA (n, F) float32 / uint8."""
    import torch
    p = _cuda(packed, torch.uint8)
    n, F = int(p.shape[0]), int(t_len)
    if int(p.shape[1]) * 8 < F:
        raise ValueError("%d packed bytes per row cannot hold %d targets" % (int(p.shape[1]), F))
    out = torch.empty((n, F), dtype=torch.float32 if dtype == "float32" else torch.uint8, device=p.device)
    _lib.check(_lib.load().sb_unpack_targets(_lib.ptr(p), n, F, _lib.SB_F32 if dtype == "float32" else _lib.SB_U8,
                                             _lib.ptr(out), _lib.stream_ptr()), "sb_unpack_targets")
    return out


def pack_samples_device(sequences=None, targets=None):
    """The inverse (what ``sampler_write_to_h5.py --compress`` stores): CUDA / numpy float (n, L, 4) and / or (n, F)
    -> (uint8 (n, ceil(L/8), 4), uint8 (n, ceil(F/8))) CUDA tensors (``None`` for an absent input)."""
    import torch
    ps = pt = None
    n = L = F = 0
    xs = ys = None
    if sequences is not None:
        xs = _cuda(sequences, torch.float32)
        n, L = int(xs.shape[0]), int(xs.shape[1])
        ps = torch.empty((n, (L + 7) // 8, 4), dtype=torch.uint8, device=xs.device)
    if targets is not None:
        ys = _cuda(targets, torch.float32)
        n, F = int(ys.shape[0]), int(ys.shape[1])
        pt = torch.empty((n, (F + 7) // 8), dtype=torch.uint8, device=ys.device)
    _lib.check(_lib.load().sb_pack_samples(_lib.ptr(xs), _lib.ptr(ys), n, max(L, 1), max(F, 1), _lib.ptr(ps),
                                           _lib.ptr(pt), _lib.stream_ptr()), "sb_pack_samples")
    return ps, pt


def unpackbits_sequence(sequence, s_len):
    """dataloader.py:129-138, same arguments and return value (float64 numpy array, unknown rows = 1/4); a single
    (L/8, 4) sample or a batch."""
    a = np.asarray(sequence)
    single = a.ndim == 2
    out = unpack_sequences_device(a[None] if single else a, s_len).cpu().numpy().astype(np.float64)
    return out[0] if single else out


def unpackbits_targets(targets, t_len):
    """dataloader.py:140-146."""
    a = np.asarray(targets)
    single = a.ndim == 1
    out = unpack_targets_device(a[None] if single else a, t_len).cpu().numpy().astype(np.float64)
    return out[0] if single else out


class H5Dataset(object):
    """``_H5Dataset`` (dataloader.py:149-265) with the packed arrays resident in HBM.  ``file_path``: an HDF5 file (needs
    h5py) or a mapping with the same keys.  ``unpackbits`` implies both ``unpackbits_seq`` and ``unpackbits_tgt``;
    without them the arrays are taken as stored (float sequences, uint8 targets).  ``use_seq_len`` / ``shift`` select the
    centre window exactly as :236-246."""

    def __init__(self, file_path, in_memory=True, unpackbits=False, unpackbits_seq=False, unpackbits_tgt=False,
                 sequence_key="sequences", targets_key="targets", use_seq_len=None, shift=None):
        self.file_path = file_path
        self.in_memory = in_memory
        self.unpackbits_seq = bool(unpackbits or unpackbits_seq)
        self.unpackbits_tgt = bool(unpackbits or unpackbits_tgt)
        self.use_seq_len, self.shift = use_seq_len, shift
        self._sequence_key, self._targets_key = sequence_key, targets_key
        self._initialized = False

    def _init(self):
        if self._initialized:
            return
        import torch
        db = self.file_path
        if isinstance(db, str):
            try:
                import h5py
            except ImportError as e:
                raise ImportError("reading %r needs the h5py package, which is not installed in this environment; pass "
                                  "a mapping with the keys %r, %r (+ '_length' entries) instead"
                                  % (db, self._sequence_key, self._targets_key)) from e
            db = h5py.File(db, "r")
        seqs = np.asarray(db[self._sequence_key])
        tgts = np.asarray(db[self._targets_key])
        self.s_len = int(np.asarray(db["{0}_length".format(self._sequence_key)])[()]) if self.unpackbits_seq \
            else int(seqs.shape[1])
        self.t_len = int(np.asarray(db["{0}_length".format(self._targets_key)])[()]) if self.unpackbits_tgt \
            else int(tgts.shape[1])
        self.sequences = torch.from_numpy(np.ascontiguousarray(seqs)).cuda()      # one upload; batches gather on the device
        self.targets = torch.from_numpy(np.ascontiguousarray(tgts)).cuda()
        self._seq_start, self._seq_end = 0, self.s_len
        if self.use_seq_len is not None:                                          # :238-246
            mid = self.s_len // 2
            self._seq_start = int(mid - np.ceil(self.use_seq_len / 2))
            self._seq_end = mid + self.use_seq_len // 2
            if self.shift is not None:
                self._seq_start += self.shift
                self._seq_end += self.shift
        self._initialized = True

    def __len__(self):
        self._init()
        return int(self.sequences.shape[0])

    def batch(self, indices, seq_dtype="float32", tgt_dtype="float32"):
        """Samples ``indices`` -> (CUDA (B, L, 4), CUDA (B, F)) ready for the training step."""
        import torch
        self._init()
        idx = torch.as_tensor(np.asarray(indices, dtype=np.int64) % len(self), device=self.sequences.device)
        seqs, tgts = self.sequences.index_select(0, idx), self.targets.index_select(0, idx)
        if self.unpackbits_seq:
            seqs = unpack_sequences_device(seqs, self.s_len, self._seq_start, self._seq_end - self._seq_start, seq_dtype)
        else:
            seqs = seqs[:, self._seq_start:self._seq_end].to(torch.float32 if seq_dtype == "float32" else torch.bfloat16)
        if self.unpackbits_tgt:
            tgts = unpack_targets_device(tgts, self.t_len, tgt_dtype)
        else:
            tgts = tgts.to(torch.float32 if tgt_dtype == "float32" else torch.uint8)
        return seqs, tgts

    def __getitem__(self, index):
        """One sample as the reference returns it: (float32 tensor (L, 4), targets tensor (F,)) on the host."""
        s, t = self.batch([index] if isinstance(index, (int, np.integer)) else index)
        s, t = s.cpu(), t.cpu().double()
        return (s[0], t[0]) if isinstance(index, (int, np.integer)) else (s, t)


class H5DataLoader(object):
    """``H5DataLoader`` (dataloader.py:268-385): iterates the dataset in batches — shuffled with a ``torch.Generator``
    seeded like the reference's, restricted by ``use_subset`` (int, (start, end) or list) — and yields CUDA tensors.
    ``num_workers`` is accepted and unused: there is no per-sample host work left to parallelise."""

    def __init__(self, dataset, num_workers=1, use_subset=None, batch_size=1, seed=436, shuffle=True, drop_last=False):
        self.dataset, self.batch_size, self.seed = dataset, int(batch_size), seed
        if isinstance(use_subset, int):
            use_subset = list(range(use_subset))
        elif isinstance(use_subset, tuple) and len(use_subset) == 2:
            use_subset = list(range(use_subset[0], use_subset[1]))
        self._subset = None if use_subset is None else np.asarray(use_subset, dtype=np.int64)
        self.shuffle = bool(shuffle or use_subset is not None)     # SubsetRandomSampler shuffles (:372-378)
        self.drop_last = drop_last
        self._epoch = 0

    def __len__(self):
        n = len(self.dataset) if self._subset is None else len(self._subset)
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def __iter__(self):
        import torch
        order = np.arange(len(self.dataset)) if self._subset is None else self._subset
        if self.shuffle:
            g = torch.Generator()
            g.manual_seed(self.seed + self._epoch)
            order = order[torch.randperm(len(order), generator=g).numpy()]
        self._epoch += 1
        for b in range(len(self)):
            yield self.dataset.batch(order[b * self.batch_size:(b + 1) * self.batch_size])
