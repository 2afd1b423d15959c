"""``Genome`` — drop-in for ``selene_sdk.sequences.Genome`` (reference: selene_sdk/sequences/genome.py)
whose encodings are produced by the sm_100a kernels behind ``libselene_b200.so``.

Same public surface as the reference class (genome.py:170-579): class-level ``BASES_ARR`` /
``BASE_TO_INDEX`` / ``INDEX_TO_BASE`` / ``COMPLEMENTARY_BASE_DICT`` / ``UNK_BASE`` with
``update_bases_order``; lazy initialisation so instances stay picklable; ``get_chrs``,
``get_chr_lens``, ``coords_in_bounds``, ``get_sequence_from_coords``, ``get_encoding_from_coords``,
``get_encoding_from_coords_check_unk``, ``sequence_to_encoding``, ``encoding_to_sequence``.

What changed underneath: the FASTA is read once (using its ``.fai`` index, no pyfaidx) into a device
image of 2 bits/base + an unknown-character bitmask + an uppercase-'N' bitmask; windows are decoded
on the GPU.  Batched entry points (``encode_windows``, ``window_codes``) return CUDA tensors and are
what the samplers / AnalyzeSequences fast paths use; the single-window reference methods call the
same kernels with n = 1 and copy the result back as the numpy array the reference returns.
"""
import ctypes as C
import gzip
import os
from functools import wraps

import numpy as np

from .. import _lib

# byte -> canonical code (A=0 C=1 G=2 T=3), 4 = anything else
_CODE_LUT = np.full(256, 4, dtype=np.uint8)
for _i, _b in enumerate("ACGT"):
    _CODE_LUT[ord(_b)] = _i
    _CODE_LUT[ord(_b.lower())] = _i
_COMP_LUT = np.arange(256, dtype=np.uint8)
for _a, _b in zip("ACGTacgt", "TGCAtgca"):
    _COMP_LUT[ord(_a)] = ord(_b)

_IMAGE_ALIGN = 64  # chromosome offsets in the image are multiples of 64 bases (16 packed bytes)


def codes_of_string(sequence):
    """Canonical base codes (uint8, 0..4) of a python string."""
    raw = np.frombuffer(sequence.encode("latin-1", "replace"), dtype=np.uint8)
    return _CODE_LUT[raw]


def _check_coords(len_chrs, chrom, start, end, pad=False, blacklist=None):
    """genome.py:51-92."""
    return chrom in len_chrs and \
        start < len_chrs[chrom] and \
        start < end and \
        end > 0 and \
        (start >= 0 if not pad else True) and \
        (end <= len_chrs[chrom] if not pad else True) and \
        (blacklist is None or not blacklist.overlaps(chrom, start, end))


class _Blacklist(object):
    """Host-side replacement of the tabix blacklist query (genome.py:17-48): regions overlapping
    [start, end) exclude the query."""

    def __init__(self, path):
        opener = gzip.open if path.endswith(".gz") else open
        by_chrom = {}
        with opener(path, "rt") as fh:
            for line in fh:
                if not line.strip() or line.startswith(("#", "track", "browser")):
                    continue
                cols = line.split("\t")
                by_chrom.setdefault(cols[0], []).append((int(cols[1]), int(cols[2])))
        self._starts, self._cmax = {}, {}
        for chrom, ivs in by_chrom.items():
            ivs.sort()
            arr = np.array(ivs, dtype=np.int64)
            self._starts[chrom] = arr[:, 0]
            self._cmax[chrom] = np.maximum.accumulate(arr[:, 1])

    def overlaps(self, chrom, start, end):
        if chrom not in self._starts:
            return False
        # intervals with start < end; among them any whose end > start (running max is monotone)
        hi = int(np.searchsorted(self._starts[chrom], end, side="left"))
        return hi > 0 and self._cmax[chrom][hi - 1] > start


def read_fasta(path):
    """Returns [(name, uint8 array of sequence bytes)] in file order, using ``path + '.fai'`` when
    present (name, length, offset, line bases, line width) and a plain parse otherwise."""
    out = []
    fai = path + ".fai"
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as fh:
        data = fh.read()
    buf = np.frombuffer(data, dtype=np.uint8)
    if os.path.exists(fai):
        with open(fai) as fh:
            for line in fh:
                cols = line.rstrip("\n").split("\t")
                if len(cols) < 5:
                    continue
                name, length, offset, lb, lw = cols[0], int(cols[1]), int(cols[2]), int(cols[3]), int(cols[4])
                if length == 0:
                    out.append((name, np.zeros(0, dtype=np.uint8)))
                    continue
                n_lines = (length + lb - 1) // lb
                span = (n_lines - 1) * lw + (length - (n_lines - 1) * lb)
                chunk = buf[offset:offset + span]
                if lw == lb or n_lines == 1:
                    seq = chunk[:length]
                else:
                    full = (n_lines - 1)
                    body = chunk[:full * lw].reshape(full, lw)[:, :lb].reshape(-1)
                    seq = np.concatenate([body, chunk[full * lw:]])
                out.append((name, np.ascontiguousarray(seq[:length])))
        return out
    name, parts = None, []
    for line in data.split(b"\n"):
        if line.startswith(b">"):
            if name is not None:
                out.append((name, np.frombuffer(b"".join(parts), dtype=np.uint8)))
            name, parts = line[1:].split()[0].decode(), []
        elif name is not None:
            parts.append(line.strip())
    if name is not None:
        out.append((name, np.frombuffer(b"".join(parts), dtype=np.uint8)))
    return out


def pack_image(seqs):
    """seqs: list of uint8 byte arrays.  Returns (packed2bit, unk_mask, upn_mask, n_image, offsets)."""
    offsets, total = [], 0
    for s in seqs:
        offsets.append(total)
        total += (len(s) + _IMAGE_ALIGN - 1) // _IMAGE_ALIGN * _IMAGE_ALIGN
    total = max(total, _IMAGE_ALIGN)
    codes = np.zeros(total, dtype=np.uint8)
    unk = np.zeros(total, dtype=np.uint8)
    upn = np.zeros(total, dtype=np.uint8)
    for s, off in zip(seqs, offsets):
        c = _CODE_LUT[s]
        u = c == 4
        codes[off:off + len(s)] = np.where(u, 0, c)
        unk[off:off + len(s)] = u
        upn[off:off + len(s)] = s == ord("N")
    q = codes.reshape(-1, 4)
    packed = (q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)).astype(np.uint8)
    return (np.ascontiguousarray(packed), np.packbits(unk, bitorder="little"),
            np.packbits(upn, bitorder="little"), total, np.array(offsets, dtype=np.int64))


class Genome(object):
    """See the module docstring.  Parameters as the reference (genome.py:170-211):
    ``input_path``, ``blacklist_regions`` (a BED/BED.gz path; the reference's bundled "hg19"/"hg38"
    files are data, not code, and are looked up under ``$SELENE_B200_DATA`` if requested),
    ``bases_order``, ``init_unpicklable``; plus ``device`` (CUDA ordinal, default current)."""

    BASES_ARR = ['A', 'C', 'G', 'T']
    BASE_TO_INDEX = {
        'A': 0, 'C': 1, 'G': 2, 'T': 3,
        'a': 0, 'c': 1, 'g': 2, 't': 3,
    }
    INDEX_TO_BASE = {0: 'A', 1: 'C', 2: 'G', 3: 'T'}
    COMPLEMENTARY_BASE_DICT = {
        'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A', 'N': 'N',
        'a': 'T', 'c': 'G', 'g': 'C', 't': 'A', 'n': 'N'
    }
    UNK_BASE = "N"

    def __init__(self, input_path, blacklist_regions=None, bases_order=None,
                 init_unpicklable=False, device=None):
        self.input_path = input_path
        self.blacklist_regions = blacklist_regions
        self.device = device
        self._initialized = False
        if bases_order is not None:
            bases = [str.upper(b) for b in bases_order]
            self.BASES_ARR = bases
            lc_bases = [str.lower(b) for b in bases]
            self.BASE_TO_INDEX = {
                **{b: ix for (ix, b) in enumerate(bases)},
                **{b: ix for (ix, b) in enumerate(lc_bases)}}
            self.INDEX_TO_BASE = {ix: b for (ix, b) in enumerate(bases)}
            self.update_bases_order(bases)
        if init_unpicklable:
            self._unpicklable_init()

    # -- class-level alphabet state (genome.py:279-286) -------------------------------------------
    @classmethod
    def update_bases_order(cls, bases):
        cls.BASES_ARR = bases
        lc_bases = [str.lower(b) for b in bases]
        cls.BASE_TO_INDEX = {
            **{b: ix for (ix, b) in enumerate(bases)},
            **{b: ix for (ix, b) in enumerate(lc_bases)}}
        cls.INDEX_TO_BASE = {ix: b for (ix, b) in enumerate(bases)}

    @classmethod
    def perm(cls):
        """Channel of each canonical base (A, C, G, T) under the current ``BASES_ARR``."""
        return [cls.BASE_TO_INDEX[b] for b in "ACGT"]

    # -- lazy initialisation (genome.py:288-316) --------------------------------------------------
    def __getstate__(self):
        state = dict(self.__dict__)
        for k in ("_handle", "_seqs", "_chrom_index", "chrs", "len_chrs", "_blacklist", "_chrom_off", "_unk_prefix"):
            state.pop(k, None)
        state["_initialized"] = False
        return state

    @classmethod
    def from_sequences(cls, names, sequences, device=None):
        """Genome over in-memory sequences (uint8 ASCII arrays or str) instead of a FASTA file."""
        g = cls(input_path=None, device=device)
        g._memory_records = [(n, np.frombuffer(s.encode("latin-1"), dtype=np.uint8) if isinstance(s, str)
                              else np.asarray(s, dtype=np.uint8)) for n, s in zip(names, sequences)]
        g._unpicklable_init()
        return g

    def _unpicklable_init(self):
        if self._initialized:
            return
        import torch
        recs = getattr(self, "_memory_records", None)
        if recs is None:
            recs = read_fasta(self.input_path)
        self._seqs = {name: seq for name, seq in recs}
        self.chrs = sorted(self._seqs.keys())
        self.len_chrs = {c: len(self._seqs[c]) for c in self.chrs}
        self._chrom_index = {c: i for i, c in enumerate(self.chrs)}
        packed, unk, upn, n_image, offs = pack_image([self._seqs[c] for c in self.chrs])
        lens = np.array([self.len_chrs[c] for c in self.chrs], dtype=np.int64)
        self._chrom_off = offs
        self._blacklist = None
        if self.blacklist_regions is not None:
            path = self.blacklist_regions
            if path in ("hg19", "hg38"):
                root = os.environ.get("SELENE_B200_DATA", "")
                fname = ("hg19_blacklist_ENCFF001TDO.bed.gz" if path == "hg19" else "hg38.blacklist.bed.gz")
                path = os.path.join(root, fname)
                if not os.path.exists(path):
                    raise FileNotFoundError(
                        "blacklist '%s' requested but %s not found; set SELENE_B200_DATA to the "
                        "directory holding the ENCODE blacklist files" % (self.blacklist_regions, path))
            self._blacklist = _Blacklist(path)
        lib = _lib.load()
        if self.device is None:
            self.device = torch.cuda.current_device()
        handle = C.c_void_p()
        _lib.check(lib.sb_genome_create(_lib.ptr(packed), _lib.ptr(unk), _lib.ptr(upn), n_image,
                                        _lib.ptr(offs), _lib.ptr(lens), len(self.chrs), int(self.device),
                                        C.byref(handle)), "sb_genome_create")
        self._handle = handle
        self._initialized = True

    def init(func):
        @wraps(func)
        def dfunc(self, *args, **kwargs):
            self._unpicklable_init()
            return func(self, *args, **kwargs)
        return dfunc

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            try:
                _lib.load().sb_genome_destroy(h)
            except Exception:
                pass
            self._handle = None

    # -- reference API ------------------------------------------------------------------------------
    @init
    def get_chrs(self):
        return self.chrs

    @init
    def get_chr_lens(self):
        return [(k, self.len_chrs[k]) for k in self.get_chrs()]

    @init
    def coords_in_bounds(self, chrom, start, end):
        return _check_coords(self.len_chrs, chrom, start, end, blacklist=self._blacklist)

    @init
    def get_sequence_from_coords(self, chrom, start, end, strand='+', pad=False):
        """genome.py:96-166 on the host copy of the sequence bytes (string API parity)."""
        if not _check_coords(self.len_chrs, chrom, start, end, pad=pad, blacklist=self._blacklist):
            return ""
        if strand != '+' and strand != '-' and strand != '.':
            raise ValueError(
                "Strand must be one of '+', '-', or '.'. Input was {0}".format(strand))
        end_pad = start_pad = 0
        if end > self.len_chrs[chrom]:
            end_pad = end - self.len_chrs[chrom]
            end = self.len_chrs[chrom]
        if start < 0:
            start_pad = -1 * start
            start = 0
        inner = self._seqs[chrom][start:end]
        if strand == '-':
            inner = _COMP_LUT[inner][::-1]
        return (self.UNK_BASE * start_pad + inner.tobytes().decode("latin-1") +
                self.UNK_BASE * end_pad)

    @init
    def get_encoding_from_coords(self, chrom, start, end, strand='+', pad=False):
        return self.get_encoding_from_coords_check_unk(chrom, start, end, strand=strand, pad=pad)[0]

    @init
    def get_encoding_from_coords_check_unk(self, chrom, start, end, strand='+', pad=False):
        """genome.py:488-542.  Invalid coordinates -> ``(0, 4)`` array, False."""
        if not _check_coords(self.len_chrs, chrom, start, end, pad=pad, blacklist=self._blacklist):
            return np.zeros((0, len(self.BASES_ARR)), dtype=np.float32), False
        if strand != '+' and strand != '-' and strand != '.':
            raise ValueError(
                "Strand must be one of '+', '-', or '.'. Input was {0}".format(strand))
        enc, flags = self.encode_windows([chrom], [start], end - start, strands=[strand])
        return enc[0].cpu().numpy(), bool(int(flags[0]) & _lib.SB_FLAG_HAS_UPPER_N)

    @classmethod
    def sequence_to_encoding(cls, sequence):
        """genome.py:545-560 -> _sequence.pyx:11-25, on the GPU (sb_encode_codes)."""
        import torch
        if len(sequence) == 0:
            return np.zeros((0, len(cls.BASES_ARR)), dtype=np.float32)
        codes = torch.from_numpy(codes_of_string(sequence).copy()).cuda()
        out = torch.empty((len(sequence), 4), dtype=torch.float32, device=codes.device)
        _lib.check(_lib.load().sb_encode_codes(_lib.ptr(codes), 1, len(sequence), _lib.perm_arg(cls.perm()),
                                               _lib.SB_F32, _lib.ptr(out), _lib.stream_ptr()),
                   "sb_encode_codes")
        return out.cpu().numpy()

    @classmethod
    def encoding_to_sequence(cls, encoding):
        """sequence.py:44-85."""
        sequence = []
        for row in np.asarray(encoding):
            base_pos = -1
            for index, val in enumerate(row):
                if val == 1:
                    base_pos = index
                    break
            sequence.append(cls.UNK_BASE if base_pos == -1 else cls.BASES_ARR[base_pos])
        return "".join(sequence)

    # -- host-side unknown-base counts (sampler rejection test without a device round trip) ----------
    @init
    def count_unknown(self, chrom, start, end):
        """Number of positions in [start, end) of ``chrom`` whose character is outside ACGTacgt (what
        ``sequence_to_encoding`` turns into 0.25 rows); positions outside the chromosome count too."""
        n = self.len_chrs[chrom]
        out = max(0, -start) + max(0, end - n)
        start, end = max(start, 0), min(end, n)
        if end <= start:
            return out
        if not hasattr(self, "_unk_prefix"):
            self._unk_prefix = {}
        if chrom not in self._unk_prefix:
            unk = _CODE_LUT[self._seqs[chrom]] == 4
            bits = np.packbits(unk, bitorder="little")
            pop = np.unpackbits(bits.reshape(-1, 1), axis=1).sum(axis=1, dtype=np.int64)
            self._unk_prefix[chrom] = (bits, np.concatenate([[0], np.cumsum(pop)]))
        bits, prefix = self._unk_prefix[chrom]

        def upto(x):  # unknown positions in [0, x)
            b, r = x >> 3, x & 7
            v = int(prefix[b])
            if r:
                v += bin(int(bits[b]) & ((1 << r) - 1)).count("1")
            return v
        return out + upto(end) - upto(start)

    # -- batched device API -------------------------------------------------------------------------
    def _window_args(self, chroms, starts, strands):
        import torch
        dev = torch.device("cuda", int(self.device))
        if isinstance(chroms, torch.Tensor):
            chrom_t = chroms.to(device=dev, dtype=torch.int32)
        else:
            idx = np.fromiter((self._chrom_index[c] if isinstance(c, str) else int(c) for c in chroms),
                              dtype=np.int32)
            chrom_t = torch.from_numpy(idx).to(dev)
        if isinstance(starts, torch.Tensor):
            start_t = starts.to(device=dev, dtype=torch.int64)
        else:
            start_t = torch.from_numpy(np.asarray(starts, dtype=np.int64)).to(dev)
        rc_t = None
        if strands is not None:
            if isinstance(strands, torch.Tensor):
                rc_t = strands.to(device=dev, dtype=torch.uint8)
            else:
                rc = np.fromiter(((1 if s in ('-', 1, True) else 0) for s in strands), dtype=np.uint8)
                rc_t = torch.from_numpy(rc).to(dev)
        return chrom_t, start_t, rc_t

    @init
    def chrom_indices(self, chroms):
        return np.fromiter((self._chrom_index[c] for c in chroms), dtype=np.int32)

    @init
    def encode_windows(self, chroms, starts, length, strands=None, dtype="float32"):
        """One-hot windows ``[start, start+length)`` -> (CUDA tensor (n, length, 4), uint8 flags (n)).
        Out-of-range parts are N-filled as with ``pad=True``; validity checks are the caller's."""
        import torch
        chrom_t, start_t, rc_t = self._window_args(chroms, starts, strands)
        n = chrom_t.numel()
        tdt = torch.float32 if dtype == "float32" else torch.bfloat16
        out = torch.empty((n, length, 4), dtype=tdt, device=chrom_t.device)
        flags = torch.empty((n,), dtype=torch.uint8, device=chrom_t.device)
        _lib.check(_lib.load().sb_encode_windows(
            self._handle, _lib.ptr(chrom_t), _lib.ptr(start_t), _lib.ptr(rc_t), n, int(length),
            _lib.perm_arg(self.perm()), _lib.SB_F32 if dtype == "float32" else _lib.SB_BF16,
            _lib.ptr(out), _lib.ptr(flags), _lib.stream_ptr()), "sb_encode_windows")
        return out, flags

    @init
    def window_codes(self, chroms, starts, length, strands=None):
        """Canonical base codes of windows -> (uint8 CUDA tensor (n, length), uint8 flags (n))."""
        import torch
        chrom_t, start_t, rc_t = self._window_args(chroms, starts, strands)
        n = chrom_t.numel()
        out = torch.empty((n, length), dtype=torch.uint8, device=chrom_t.device)
        flags = torch.empty((n,), dtype=torch.uint8, device=chrom_t.device)
        _lib.check(_lib.load().sb_genome_codes(
            self._handle, _lib.ptr(chrom_t), _lib.ptr(start_t), _lib.ptr(rc_t), n, int(length),
            _lib.ptr(out), _lib.ptr(flags), _lib.stream_ptr()), "sb_genome_codes")
        return out, flags

    @init
    def snv_pairs(self, chroms, starts, ref_codes, alt_codes, length, var_row, strands=None,
                  dtype="float32", want_onehot=True, want_codes=False):
        """Batched body of the VEP per-variant loop (model_predict.py:1054-1110) for SNVs.
        Returns dict(ref, alt, codes, ref_match, flags) of CUDA tensors (absent ones None)."""
        import torch
        chrom_t, start_t, rc_t = self._window_args(chroms, starts, strands)
        dev = chrom_t.device
        n = chrom_t.numel()

        def as_u8(x):
            if isinstance(x, torch.Tensor):
                return x.to(device=dev, dtype=torch.uint8)
            return torch.from_numpy(np.asarray(x, dtype=np.uint8)).to(dev)
        ref_t, alt_t = as_u8(ref_codes), as_u8(alt_codes)
        tdt = torch.float32 if dtype == "float32" else torch.bfloat16
        ref_o = torch.empty((n, length, 4), dtype=tdt, device=dev) if want_onehot else None
        alt_o = torch.empty((n, length, 4), dtype=tdt, device=dev) if want_onehot else None
        codes = torch.empty((n, length), dtype=torch.uint8, device=dev) if want_codes else None
        match = torch.empty((n,), dtype=torch.uint8, device=dev)
        flags = torch.empty((n,), dtype=torch.uint8, device=dev)
        _lib.check(_lib.load().sb_snv_pairs(
            self._handle, _lib.ptr(chrom_t), _lib.ptr(start_t), _lib.ptr(ref_t), _lib.ptr(alt_t),
            _lib.ptr(rc_t), n, int(length), int(var_row), _lib.perm_arg(self.perm()),
            _lib.SB_F32 if dtype == "float32" else _lib.SB_BF16, _lib.ptr(ref_o), _lib.ptr(alt_o),
            _lib.ptr(codes), _lib.ptr(match), _lib.ptr(flags), _lib.stream_ptr()), "sb_snv_pairs")
        return dict(ref=ref_o, alt=alt_o, codes=codes, ref_match=match, flags=flags)
