"""``GenomicFeatures`` — drop-in for ``selene_sdk.targets.GenomicFeatures`` (reference:
selene_sdk/targets/genomic_features.py:196-418) labelling query bins with kernel (b)
(``sb_label_bins``) instead of a tabix query + ``_fast_get_feature_data``
(selene_sdk/targets/_genomic_features.pyx:12-41).

The BED(.gz) file is read once (plain gzip; the ``.tbi`` index is not needed) into sorted interval
arrays in HBM.  ``get_feature_data`` keeps the reference's signature, dtypes and corner cases:
int64 labels with thresholds, float64 zeros for an unknown contig, float64 0/1 in the no-threshold
mode, ``KeyError`` when a returned row names a feature that is not in ``features``.
"""
import ctypes as C
import gzip
import types
from functools import wraps

import numpy as np

from .. import _lib


def _is_positive_row(start, end, feature_start, feature_end, threshold):
    """genomic_features.py:63-101."""
    overlap_start = max(feature_start, start)
    overlap_end = min(feature_end, end)
    min_overlap_needed = int((end - start) * threshold - 1)
    if min_overlap_needed < 0:
        min_overlap_needed = 0
    return overlap_end - overlap_start > min_overlap_needed


def _define_feature_thresholds(feature_thresholds, features):
    """(dict feature -> threshold, float32 vector in feature order) from the four accepted forms of
    ``feature_thresholds`` (genomic_features.py:144-193): a number in (0, 1] for every feature; a dict with a
    ``"default"`` entry and per-feature overrides; a function of the feature name; anything else leaves the
    thresholds at zero with an empty dict, as upstream."""
    if isinstance(feature_thresholds, (float, int)) and 0 < feature_thresholds <= 1:
        per_feature = [feature_thresholds] * len(features)
    elif isinstance(feature_thresholds, dict):
        per_feature = [feature_thresholds.get(f, feature_thresholds["default"]) for f in features]
    elif isinstance(feature_thresholds, types.FunctionType):
        per_feature = [feature_thresholds(f) for f in features]
    else:
        return {}, np.zeros(len(features), dtype=np.float32)
    return dict(zip(features, per_feature)), np.array(per_feature, dtype=np.float64).astype(np.float32)


def integer_thresholds(thresholds_vec, query_length):
    """The integer threshold of _genomic_features.pyx:39-40, computed with exactly that numpy
    expression (float32 arithmetic, clip, truncate) so the kernel compares integers."""
    t = (np.asarray(thresholds_vec, dtype=np.float32) * int(query_length) - 1).clip(min=0)
    return t.astype(np.int64)


def read_bed(path):
    """-> (chrom names array, start int64, end int64, list of remaining column tuples)."""
    opener = gzip.open if path.endswith(".gz") else open
    chroms, starts, ends, rest = [], [], [], []
    with opener(path, "rt") as fh:
        for line in fh:
            line = line.rstrip("\n")
            if not line or line.startswith("#"):
                continue
            cols = line.split("\t")
            chroms.append(cols[0])
            starts.append(int(cols[1]))
            ends.append(int(cols[2]))
            rest.append(cols[3:])
    return chroms, np.array(starts, dtype=np.int64), np.array(ends, dtype=np.int64), rest


class GenomicFeatures(object):
    """See the module docstring; constructor arguments as genomic_features.py:277-296."""

    def __init__(self, input_path, features, feature_thresholds=None, init_unpicklable=False,
                 device=None):
        self.input_path = input_path
        self.n_features = len(features)
        self.feature_index_dict = dict([(feat, index) for index, feat in enumerate(features)])
        self.index_feature_dict = dict(list(enumerate(features)))
        if feature_thresholds is None:
            self.feature_thresholds = None
            self._feature_thresholds_vec = None
        else:
            self.feature_thresholds, self._feature_thresholds_vec = \
                _define_feature_thresholds(feature_thresholds, features)
        self.device = device
        self._initialized = False
        if init_unpicklable:
            self._unpicklable_init()

    def __getstate__(self):
        state = dict(self.__dict__)
        for k in ("_handle", "_chrom_index", "_bad", "_thr_cache"):
            state.pop(k, None)
        state["_initialized"] = False
        return state

    def _unpicklable_init(self):
        if self._initialized:
            return
        import torch
        chroms, starts, ends, rest = read_bed(self.input_path)
        names = sorted(set(chroms))
        self._chrom_index = {c: i for i, c in enumerate(names)}
        # threshold mode reads the feature from column 3 (_genomic_features.pyx:34); the
        # no-threshold mode from the LAST column (genomic_features.py:411)
        col = 0 if self._feature_thresholds_vec is not None else -1
        cidx = np.fromiter((self._chrom_index[c] for c in chroms), dtype=np.int32, count=len(chroms))
        fidx = np.fromiter((self.feature_index_dict.get(r[col] if r else None, -1) for r in rest),
                           dtype=np.int32, count=len(rest))
        good = fidx >= 0
        # rows naming an unknown feature raise KeyError when a query returns them (as the reference)
        self._bad = {}
        if not good.all():
            for ci in np.unique(cidx[~good]):
                sel = (~good) & (cidx == ci)
                order = np.argsort(starts[sel], kind="stable")
                bad_names = [r[col] if r else None for r, s in zip(rest, sel) if s]
                self._bad[int(ci)] = (starts[sel][order], ends[sel][order], np.maximum.accumulate(ends[sel][order]),
                                      [bad_names[k] for k in order])
        if self.device is None:
            self.device = torch.cuda.current_device()
        lib = _lib.load()
        h = C.c_void_p()
        c32 = np.ascontiguousarray(cidx[good])
        s32 = np.ascontiguousarray(starts[good].astype(np.int32))
        e32 = np.ascontiguousarray(ends[good].astype(np.int32))
        f32 = np.ascontiguousarray(fidx[good])
        _lib.check(lib.sb_intervals_create(_lib.ptr(c32), _lib.ptr(s32), _lib.ptr(e32), _lib.ptr(f32),
                                           int(good.sum()), self.n_features, max(len(names), 1),
                                           int(self.device), C.byref(h)), "sb_intervals_create")
        self._handle = h
        self._thr_cache = {}
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
                _lib.load().sb_intervals_destroy(h)
            except Exception:
                pass
            self._handle = None

    def _check_bad(self, ci, start, end):
        """Raises ``KeyError(feature)`` for the first row (in start order, as tabix returns them) that overlaps the
        query and names a feature outside ``feature_index_dict`` — what ``feature_index_dict[row[3]]`` does upstream
        (_genomic_features.pyx:34)."""
        if ci in self._bad:
            s, e, cmax, names = self._bad[ci]
            hi = int(np.searchsorted(s, end, side="left"))         # rows with start < end of the query
            if hi > 0 and cmax[hi - 1] > start:
                first = int(np.argmax(e[:hi] > start))             # first of them that ends after the query start
                raise KeyError(names[first])

    def _thr_tensor(self, query_length):
        import torch
        key = int(query_length)
        if key not in self._thr_cache:
            if self._feature_thresholds_vec is None:
                t = np.full(self.n_features, -1, dtype=np.int32)
            else:
                t = integer_thresholds(self._feature_thresholds_vec, key).astype(np.int32)
            self._thr_cache[key] = torch.from_numpy(t).to(torch.device("cuda", int(self.device)))
        return self._thr_cache[key]

    @init
    def label_bins(self, chroms, starts, ends, out_dtype="uint8"):
        """Batched labels: CUDA tensor (n, n_features).  All bins must have the same length when
        thresholds are used (the integer threshold depends on it).  Unknown contigs give zeros."""
        import torch
        dev = torch.device("cuda", int(self.device))
        starts = np.asarray(starts, dtype=np.int64)
        ends = np.asarray(ends, dtype=np.int64)
        n = len(starts)
        qlens = np.unique(ends - starts) if n else np.array([1])
        if self._feature_thresholds_vec is not None and len(qlens) != 1:
            raise ValueError("label_bins needs equal-length bins when thresholds are set")
        ci = np.fromiter((self._chrom_index.get(c, -1) if isinstance(c, str) else int(c) for c in chroms),
                         dtype=np.int32, count=n)
        if self._bad:
            for i in range(n):
                self._check_bad(int(ci[i]), int(starts[i]), int(ends[i]))
        tdt = {"uint8": (torch.uint8, _lib.SB_U8), "float32": (torch.float32, _lib.SB_F32),
               "int64": (torch.int64, _lib.SB_I64), "float64": (torch.float64, _lib.SB_F64)}[out_dtype]
        out = torch.empty((n, self.n_features), dtype=tdt[0], device=dev)
        if n == 0:
            return out
        c_t = torch.from_numpy(ci).to(dev)
        s_t = torch.from_numpy(starts.astype(np.int32)).to(dev)
        e_t = torch.from_numpy(ends.astype(np.int32)).to(dev)
        _lib.check(_lib.load().sb_label_bins(self._handle, _lib.ptr(c_t), _lib.ptr(s_t), _lib.ptr(e_t), n,
                                             _lib.ptr(self._thr_tensor(int(qlens[0]))), tdt[1],
                                             _lib.ptr(out), _lib.stream_ptr()), "sb_label_bins")
        return out

    @init
    def get_feature_data(self, chrom, start, end, strand=None):
        """genomic_features.py:377-418."""
        if strand == '+' or strand == '-':
            # genomic_features.py:343-344 filters rows by column 3 == strand and _genomic_features.pyx:34
            # then reads the feature from that same column: broken upstream, not emulated.
            raise NotImplementedError("strand-specific targets are broken in the reference; unsupported")
        if chrom not in self._chrom_index:
            return np.zeros((self.n_features,))  # TabixError branch: float64 zeros
        if self._feature_thresholds_vec is None:
            return self.label_bins([chrom], [start], [end], out_dtype="float64")[0].cpu().numpy()
        return self.label_bins([chrom], [start], [end], out_dtype="int64")[0].cpu().numpy()

    @init
    def is_positive(self, chrom, start, end, strand=None):
        """genomic_features.py:350-375 with _any_positive_rows (:22-60): any single row whose own
        overlap exceeds int((end-start)*thr - 1) (float64 arithmetic — not the label formula)."""
        if self.feature_thresholds is None:
            raise TypeError("'NoneType' object is not subscriptable")  # what the reference does
        if chrom not in self._chrom_index:
            return False
        # evaluated with one label query per distinct threshold value using the per-row rule:
        # a row is positive iff its clipped length > t, i.e. iff that single row's coverage > t;
        # the union coverage can exceed t without any single row doing so, hence the host loop.
        rows = self._rows_overlapping(chrom, start, end)
        for fs, fe, f in rows:
            if _is_positive_row(start, end, fs, fe, self.feature_thresholds[f]):
                return True
        return False

    def _rows_overlapping(self, chrom, start, end):
        if not hasattr(self, "_host_rows"):
            chroms, starts, ends, rest = read_bed(self.input_path)
            col = 0 if self._feature_thresholds_vec is not None else -1
            by = {}
            for c, s, e, r in zip(chroms, starts, ends, rest):
                by.setdefault(c, []).append((int(s), int(e), r[col]))
            self._host_rows = by
        return [r for r in self._host_rows.get(chrom, []) if r[0] < end and r[1] > start]
