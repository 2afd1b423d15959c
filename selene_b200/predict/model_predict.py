"""``AnalyzeSequences`` — drop-in for ``selene_sdk.predict.AnalyzeSequences`` (reference:
selene_sdk/predict/model_predict.py:42-1153) for the two scoring entry points on the hot path:
``in_silico_mutagenesis`` (:662-799) and ``variant_effect_prediction`` (:952-1141).

The per-mutant / per-variant Python loops of the reference (:636-657, :1054-1129) are replaced by
whole-job device calls: windows and mutants are described by (code row, substituted position, base)
triples, the conv stack consumes them directly, and scores come out of the fused epilogue.  Batches
of ``batch_size`` exist only where the reference's behaviour depends on them (the in-place clamp of
the ISM baseline, §9.21 of SURVEY.md).  File outputs keep the reference's names, columns and number
formats (handler.py:15-114,190-235; write_ref_alt_handler.py:83-108).
"""
import math
import os
import warnings

import numpy as np

from .. import _lib
from ..nets import B200Net, layers_from_module
from ..sequences import Genome, codes_of_string
from ._in_silico_mutagenesis import ism_mutations_device
from ._variant_effect_prediction import VariantTable, read_vcf_file, build_ref_alt_strings, _get_ref_idxs  # noqa: F401

ISM_COLS = ["pos", "ref", "alt"]                                                  # model_predict.py:39
VARIANTEFFECT_COLS = ["chrom", "pos", "name", "ref", "alt", "strand", "ref_match", "contains_unk"]  # :40
_CODE = {"A": 0, "C": 1, "G": 2, "T": 3, "a": 0, "c": 1, "g": 2, "t": 3}


def _is_lua_trained_model(model):
    """selene_sdk/utils/utils.py:16-31."""
    import torch
    if hasattr(model, 'from_lua'):
        return model.from_lua
    check_model = model
    if hasattr(model, 'model'):
        check_model = model.model
    elif type(model) == torch.nn.DataParallel and hasattr(model, 'module'):
        check_model = model.module
    else:
        for m in model.modules():
            if "Conv2d" in m.__class__.__name__:
                return True
    return False


# --------------------------------------------------------------------------------------------------
# writers (same files, columns and number formats as predict_handlers/handler.py)
# --------------------------------------------------------------------------------------------------
def _scores_filepath(output_path_prefix, handler_filename):
    """handler.py:190-203."""
    output_path, filename_prefix = None, None
    if not os.path.isdir(output_path_prefix):
        output_path, filename_prefix = os.path.split(output_path_prefix)
    else:
        output_path = output_path_prefix
    if filename_prefix is not None:
        handler_filename = "{0}_{1}".format(filename_prefix, handler_filename)
    return output_path, filename_prefix, os.path.join(output_path, handler_filename)


def format_e2_rows(data, row_prefixes=None):
    """``"{:.2e}"`` text of a float32 score matrix, formatted on the device (handler.py:36-42).
    ``data``: (n, F) numpy array or CUDA tensor.  Returns (bytes of all rows as one uint8 numpy array, offsets
    (n+1,) int64): row r is ``buf[off[r]:off[r+1]]`` = the values joined by tabs + a newline, preceded by
    ``row_prefixes[r]`` (bytes, e.g. the id columns and a tab) when given."""
    import torch
    lib = _lib.load()
    t = data if isinstance(data, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32))
    t = t.to(device="cuda", dtype=torch.float32).contiguous()
    n, F = int(t.shape[0]), int(t.shape[1])
    if n == 0:
        return np.zeros(0, np.uint8), np.zeros(1, np.int64)
    sizes = torch.empty(n, dtype=torch.int64, device=t.device)
    _lib.check(lib.sb_format_e2_row_bytes(_lib.ptr(t), n, F, _lib.ptr(sizes), _lib.stream_ptr()), "sb_format_e2_row_bytes")
    pre = pre_off = None
    if row_prefixes is not None:
        if len(row_prefixes) != n:
            raise ValueError("one prefix per row is needed")
        lens = np.fromiter((len(p) for p in row_prefixes), dtype=np.int64, count=n)
        po = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(lens, out=po[1:])
        pre = torch.from_numpy(np.frombuffer(b"".join(row_prefixes), dtype=np.uint8).copy()).to(t.device)
        pre_off = torch.from_numpy(po).to(t.device)
        sizes += pre_off[1:] - pre_off[:-1]
    off = torch.zeros(n + 1, dtype=torch.int64, device=t.device)
    torch.cumsum(sizes, 0, out=off[1:])
    total = int(off[-1].item())
    buf = torch.empty(total, dtype=torch.uint8, device=t.device)
    if pre is None:
        _lib.check(lib.sb_format_e2_rows(_lib.ptr(t), n, F, _lib.ptr(off), _lib.ptr(buf), _lib.stream_ptr()),
                   "sb_format_e2_rows")
    else:
        _lib.check(lib.sb_format_e2_rows_prefixed(_lib.ptr(t), n, F, _lib.ptr(off), _lib.ptr(pre), _lib.ptr(pre_off),
                                                  _lib.ptr(buf), _lib.stream_ptr()), "sb_format_e2_rows_prefixed")
    return buf.cpu().numpy(), off.cpu().numpy()


from ._writers import (submit_tsv, write_tsv, write_hdf5, write_row_labels, prefixes_from_ids)  # noqa: E402,F401


def _write_outputs(results, ids, features, cols, output_path_prefix, output_format, defer=False, prefix=None):
    """results: ordered list of (handler_filename, prefix_override or None, array).  TSV files are written by the
    writer pool of ``_writers`` (formatted on the device, chunks written concurrently at their offsets).  ``ids``: list
    of id tuples (may be None when ``prefix`` = (bytes, offsets) of the id columns is given and the format is TSV).
    With ``defer`` the ``finish`` callables are returned instead of awaited, so the caller can score the next input
    meanwhile."""
    labels_written = False
    if output_format == "tsv" and prefix is None:
        prefix = prefixes_from_ids(ids)
    finishers = []
    for name, prefix_override, arr in results:
        opp = output_path_prefix if prefix_override is None else prefix_override
        out_dir, fprefix, sp = _scores_filepath(opp, name)
        if output_format == "tsv":
            finishers.append(submit_tsv(sp + ".tsv", cols, features, arr, prefix[0], prefix[1]))
        else:
            write_hdf5(sp + ".h5", arr)
            if not labels_written:
                lf = "row_labels.txt"
                if fprefix is not None:
                    if fprefix[-4:] in (".ref", ".alt"):
                        fprefix = fprefix[:-4]
                    lf = "{0}_{1}".format(fprefix, lf)
                write_row_labels(os.path.join(out_dir, lf), cols, ids)
                labels_written = True
    if defer:
        return finishers
    for f in finishers:
        f()      # waits, closes the file, re-raises a writer's exception
    return []


_PINNED = {}   # (name, shape, dtype) -> persistent pinned staging tensor (allocating pinned memory per call is slow)


def _to_host(tensors):
    """Device -> host as fresh numpy arrays (what the reference's handlers receive): all copies go through
    persistent pinned staging buffers and are in flight together, then one synchronisation and a host copy each
    (a pageable ``.cpu()`` of 22 MB of ISM scores runs at ~6 GB/s)."""
    import torch
    staged = {}
    for k, v in tensors.items():
        if isinstance(v, torch.Tensor) and v.is_cuda and v.numel() >= (1 << 16):
            key = (k, tuple(v.shape), v.dtype)
            h = _PINNED.get(key)
            if h is None:
                if len(_PINNED) >= 16:
                    _PINNED.clear()
                h = _PINNED[key] = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
            h.copy_(v, non_blocking=True)
            staged[k] = h
    if staged:
        torch.cuda.current_stream().synchronize()
    out = {}
    for k, v in tensors.items():
        if k in staged:
            out[k] = staged[k].numpy().copy()
        else:
            out[k] = v.cpu().numpy() if isinstance(v, torch.Tensor) else v
    return out


class AnalyzeSequences(object):
    """Score sequences and variants with a trained sequence-based model on a B200.

    Parameters follow model_predict.py:42-131: ``model`` is the ``torch.nn.Module`` (e.g.
    ``DeepSEA(1000, 919)``), ``trained_model_path`` optional weights (``state_dict`` or a checkpoint
    dict with a ``state_dict`` key), ``sequence_length``, ``features``, ``batch_size``,
    ``reference_sequence`` (a :class:`selene_b200.sequences.Genome` instance or the class).
    ``use_cuda`` must stay True: there is no CPU path.  ``precision``: ``"bf16"`` (tcgen05, default)
    or ``"fp32"`` (accuracy mode)."""

    def __init__(self, model, trained_model_path=None, sequence_length=None, features=None, batch_size=64,
                 use_cuda=True, data_parallel=False, reference_sequence=Genome, write_mem_limit=1500,
                 precision="bf16", device=None):
        import torch
        if not use_cuda:
            raise ValueError("selene_b200 has no CPU path: use_cuda must be True")
        if trained_model_path is not None:
            ck = torch.load(trained_model_path, map_location="cpu")
            state = ck["state_dict"] if isinstance(ck, dict) and "state_dict" in ck else ck
            # selene_sdk/utils/utils.py:111-126: weights are matched to the model's own keys by position
            own = model.state_dict()
            if len(own) != len(state):
                raise ValueError("The model and the weights file do not have the same number of tensors")
            model.load_state_dict({k: v for k, v in zip(own.keys(), state.values())})
        model.eval()
        self.model = model
        self.sequence_length = sequence_length
        self._start_radius = sequence_length // 2
        self._end_radius = self._start_radius
        if sequence_length % 2 != 0:
            self._start_radius += 1
        self.batch_size = batch_size
        self.features = list(features)
        self.reference_sequence = reference_sequence
        self.precision = precision
        if type(self.reference_sequence) == Genome and not self.reference_sequence._initialized:
            self.reference_sequence._unpicklable_init()
        if type(self.reference_sequence) == Genome and _is_lua_trained_model(model):
            Genome.update_bases_order(['A', 'G', 'C', 'T'])     # model_predict.py:179-181
        else:
            Genome.update_bases_order(['A', 'C', 'G', 'T'])     # :182-183
        self._write_mem_limit = write_mem_limit
        self.multi_gpu_output = "shards"     # or "gather": see variant_effect_prediction
        # single-process VEP: variants scored per block while the previous block's rows are written.  64 x 148 variants =
        # 18 944 rows = four wave-exact chunks of the DeepSEA-family nets on 148 SMs; measured file to file on one B200
        # (profiles/r02k_vep_file.md): 32 768 -> 247 k variants/s, 18 944 -> 289 k, 9 472 -> 306 k (the last block's
        # write is the part nothing overlaps)
        self.vep_block_variants = 9472
        # the weights are re-normalised for models that do it on every forward (heartenn.py:73-78)
        from ..nets import unwrap_non_strand_specific
        inner = unwrap_non_strand_specific(model)[0]
        if hasattr(inner, "_renorm") or inner.__class__.__name__ == "HeartENN":
            with torch.no_grad():
                for m in inner.modules():
                    if isinstance(m, (torch.nn.Conv1d, torch.nn.Linear)):
                        m.weight.data.renorm_(2, 0, 0.9)
        self.net = B200Net(layers_from_module(model), sequence_length, perm=Genome.perm(), device=device,
                           strand_mode=unwrap_non_strand_specific(model)[1])
        if self.net.n_outputs != len(self.features):
            raise ValueError("model predicts %d features, %d feature names given" %
                             (self.net.n_outputs, len(self.features)))

    # ----------------------------------------------------------------------------------------------
    # in-memory scoring (what the file-level methods and bench.py's e2e leg call)
    # ----------------------------------------------------------------------------------------------
    def _want(self, save_data):
        save_data = sorted(set(save_data) & set(["diffs", "abs_diffs", "logits", "predictions"]))
        if len(save_data) == 0:
            raise ValueError("'save_data' parameter must be a list that contains one of "
                             "['diffs', 'abs_diffs', 'logits', 'predictions'].")
        return save_data

    def prepare_snvs(self, chroms, positions, ref_codes, alt_codes, strands=None, pinned=None):
        """Host -> device ingest of an SNV list (the only H2D traffic of the VEP path): chromosome
        names or indices, 1-based positions, canonical base codes, optional strands.  Returns a dict of
        CUDA tensors describing the 2n interleaved (ref, alt) model rows."""
        import torch
        G = self.reference_sequence
        G._unpicklable_init()
        L = self.sequence_length
        dev = torch.device("cuda", int(G.device))
        n = len(positions)
        var_row = _get_ref_idxs(L, 1)[0]
        if len(chroms) and isinstance(chroms[0], str):
            cidx = G.chrom_indices(chroms)
        else:
            cidx = np.asarray(chroms, dtype=np.int32)
        starts = np.asarray(positions, dtype=np.int64) - self._start_radius
        ref = np.asarray(ref_codes, dtype=np.uint8)
        alt = np.asarray(alt_codes, dtype=np.uint8)
        rc = np.zeros(n, dtype=np.uint8)
        if isinstance(strands, np.ndarray) and strands.dtype != object and strands.dtype.kind in "bui":
            rc = (strands != 0).astype(np.uint8)            # already a minus-strand mask
        elif strands is not None:
            rc = np.fromiter(((1 if s in ('-', 1, True) else 0) for s in strands), dtype=np.uint8, count=n)
        # model rows (ref_i, alt_i): code row i with the VCF ref / the alt forced at the variant row;
        # on the minus strand the finished encodings are reverse-complemented (model_predict.py:1102-1110),
        # i.e. the row is mirrored and the forced base complemented
        comp = np.array([3, 2, 1, 0, 4], dtype=np.uint8)
        sub = np.empty(2 * n, dtype=np.uint8)
        sub[0::2] = np.where(rc == 1, comp[np.minimum(ref, 4)], ref)
        sub[1::2] = np.where(rc == 1, comp[np.minimum(alt, 4)], alt)
        pos = np.repeat(np.where(rc == 1, L - 1 - var_row, var_row).astype(np.int32), 2)
        src = np.repeat(np.arange(n, dtype=np.int32), 2)

        def up(a):
            t = torch.from_numpy(np.ascontiguousarray(a))
            if pinned:
                t = t.pin_memory()
            return t.to(dev, non_blocking=True)
        host = dict(chrom=cidx.astype(np.int32), start=starts, ref=ref, alt=alt, src=src, pos=pos, sub=sub)
        out = {k: up(v) for k, v in host.items()}
        out.update(n=n, var_row=var_row, rc=up(rc) if rc.any() else None,
                   h2d_bytes=int(sum(v.nbytes for v in host.values()) + (rc.nbytes if rc.any() else 0)))
        return out

    def score_prepared(self, prep, save_data=("abs_diffs", "logits"), outs=None):
        """Device-resident leg of SNV scoring: ``sb_snv_pairs`` (window codes, ref_match, flags) then the
        fused forward + score.  Returns dict of CUDA tensors."""
        want = self._want(save_data)
        G = self.reference_sequence
        res = G.snv_pairs(prep["chrom"], prep["start"], prep["ref"], prep["alt"], self.sequence_length,
                          prep["var_row"], strands=prep["rc"], want_onehot=False, want_codes=True)
        req = [s for s in want if s != "predictions"]
        if "predictions" in want:
            req += ["ref_predictions", "alt_predictions"]
        self.net.set_recurrence(self.batch_size, 2)   # ref and alt batches are predicted separately
        out = self.net.forward_score(res["codes"], prep["src"], prep["pos"], prep["sub"], 2 * prep["n"],
                                     _lib.SB_PAIR_INTERLEAVED, precision=self.precision, want=tuple(req),
                                     outs=outs)
        out["ref_match"] = res["ref_match"]
        out["flags"] = res["flags"]
        return out

    def score_snvs(self, chroms, positions, ref_codes, alt_codes, strands=None,
                   save_data=("abs_diffs", "logits"), to_host=True):
        """Scores SNVs given as host arrays.  Returns dict with the requested score arrays (n, F) plus
        ``ref_match`` and ``contains_unk`` (bool), computed as model_predict.py:1054-1124."""
        prep = self.prepare_snvs(chroms, positions, ref_codes, alt_codes, strands)
        out = self.score_prepared(prep, save_data)
        out["ref_match"] = out["ref_match"].bool()
        out["contains_unk"] = (out.pop("flags") & _lib.SB_FLAG_HAS_UPPER_N).bool()
        if to_host:
            out = _to_host(out)
        return out

    def score_code_pairs(self, ref_codes, alt_codes, save_data=("abs_diffs", "logits"), to_host=True):
        """Scores explicit (ref row, alt row) code pairs: uint8 arrays (n, L).  Used for indels / MNVs
        whose windows the host builds string-wise (_variant_effect_prediction.py:146-266)."""
        import torch
        want = self._want(save_data)
        n, L = ref_codes.shape
        inter = np.empty((2 * n, L), dtype=np.uint8)
        inter[0::2], inter[1::2] = ref_codes, alt_codes
        codes = torch.from_numpy(inter).cuda()
        self.net.set_recurrence(self.batch_size, 2)
        req = [s for s in want if s != "predictions"]
        if "predictions" in want:
            req += ["ref_predictions", "alt_predictions"]
        out = self.net.forward_score(codes, None, None, None, 2 * n, _lib.SB_PAIR_INTERLEAVED,
                                     precision=self.precision, want=tuple(req))
        if to_host:
            out = _to_host(out)
        return out

    def ism_scores(self, sequence, save_data=("abs_diffs", "logits"), start_position=0, end_position=None,
                   to_host=True, mutate_n_bases=1):
        """ISM of one (already padded/truncated, upper-cased) sequence.  Returns (mutation list, dict of arrays,
        baseline prediction (1, F)).  The mutation list is (pos int array, alt code array) for single-base ISM —
        enumerated on the device — and the reference's list of ``[(pos, base), ...]`` for ``mutate_n_bases > 1``
        (host enumeration in itertools order, _in_silico_mutagenesis.py:91-107; rows are built on the device)."""
        import torch
        want = self._want(save_data)
        if end_position is None:
            end_position = self.sequence_length
        self.net.set_recurrence(self.batch_size, 1)
        req = [s for s in want if s != "predictions"]
        if "predictions" in want:
            req.append("alt_predictions")
        if mutate_n_bases == 1:
            pos, alt, codes = ism_mutations_device(sequence, start_position, end_position, self.reference_sequence)
            muts, m = (pos, alt), int(pos.numel())
        else:
            from ._in_silico_mutagenesis import in_silico_mutagenesis_sequences
            muts = in_silico_mutagenesis_sequences(sequence, mutate_n_bases, self.reference_sequence,
                                                   start_position, end_position)
            m = len(muts)
            codes = torch.from_numpy(codes_of_string(sequence).copy()).cuda().reshape(1, -1)
            mpos = torch.from_numpy(np.array([[p for p, _ in mm] for mm in muts], dtype=np.int64).reshape(m, -1)).cuda()
            malt = torch.from_numpy(np.array([[_CODE[b] for _, b in mm] for mm in muts],
                                             dtype=np.uint8).reshape(m, -1)).cuda()
        base = self.net.forward_codes(codes, precision=self.precision)          # (1, F) baseline
        base_pair, base_index = base, None
        if "logits" in want:
            # logit_score_handler.py:118-124 clamps the shared baseline array in place during the first
            # batch, so batches after the first see the clamped baseline in EVERY handler (§9.21)
            clamped = torch.where(base == 0, torch.full_like(base, 1e-24), base)
            clamped = torch.where(clamped >= 1, torch.full_like(base, 0.999999), clamped)
            base_pair = torch.cat([base, clamped], dim=0).contiguous()
            base_index = (torch.arange(m, device=base.device) >= self.batch_size).to(torch.int32)
        if mutate_n_bases == 1:
            out = self.net.forward_score(codes, torch.zeros(m, dtype=torch.int32, device=codes.device), pos, alt, m,
                                         _lib.SB_PAIR_BASELINE, precision=self.precision, base_pred=base_pair,
                                         base_index=base_index, want=tuple(req))
        else:
            F = len(self.features)
            out = {k: torch.empty((m, F), dtype=torch.float32, device=codes.device) for k in req}
            step = max(self.batch_size, (1 << 16) // self.batch_size * self.batch_size)
            for a in range(0, m, step):
                z = min(m, a + step)
                rows = codes.repeat(z - a, 1)
                rows.scatter_(1, mpos[a:z], malt[a:z])
                self.net.forward_score(rows, None, None, None, z - a, _lib.SB_PAIR_BASELINE, precision=self.precision,
                                       base_pred=base_pair, base_index=None if base_index is None else base_index[a:z],
                                       want=tuple(req), outs={k: v[a:z] for k, v in out.items()})
        if "logits" in want:
            base = base_pair[1:2]
        if to_host:
            out = _to_host(out)
            if mutate_n_bases == 1:
                muts = (pos.cpu().numpy(), alt.cpu().numpy())
            return muts, out, base.cpu().numpy()
        return muts, out, base

    # ----------------------------------------------------------------------------------------------
    # reference entry points
    # ----------------------------------------------------------------------------------------------
    def in_silico_mutagenesis(self, sequence, save_data, output_path_prefix="ism", mutate_n_bases=1,
                              output_format="tsv", start_position=0, end_position=None, _defer_writes=False,
                              _honour_n_bases=False):
        """model_predict.py:662-799.  The reference checks ``mutate_n_bases`` but then hard-wires single-base ISM
        (:763-764), and so does this method; ``in_silico_mutagenesis_from_file`` honours it (:907-912) and calls this
        with ``_honour_n_bases``."""
        if end_position is None:
            end_position = self.sequence_length
        if start_position >= end_position:
            raise ValueError(("Starting positions must be less than the ending "
                              "positions. Found a starting position of {0} with "
                              "an ending position of {1}.").format(start_position, end_position))
        if start_position < 0:
            raise ValueError("Negative starting positions are not supported.")
        if end_position < 0:
            raise ValueError("Negative ending positions are not supported.")
        if start_position >= self.sequence_length:
            raise ValueError(("Starting positions must be less than the sequence length."
                              " Found a starting position of {0} with a sequence length "
                              "of {1}.").format(start_position, self.sequence_length))
        if end_position > self.sequence_length:
            raise ValueError(("Ending positions must be less than or equal to the sequence "
                              "length. Found an ending position of {0} with a sequence "
                              "length of {1}.").format(end_position, self.sequence_length))
        if (end_position - start_position) < mutate_n_bases:
            raise ValueError(("Fewer bases exist in the substring specified by the starting "
                              "and ending positions than need to be mutated. There are only "
                              "{0} currently, but {1} bases must be mutated at a "
                              "time").format(end_position - start_position, mutate_n_bases))
        path_dirs, _ = os.path.split(output_path_prefix)
        if path_dirs:
            os.makedirs(path_dirs, exist_ok=True)
        n = len(sequence)
        if n < self.sequence_length:
            diff = (self.sequence_length - n) / 2
            pad_l, pad_r = int(np.floor(diff)), math.ceil(diff)
            sequence = (self.reference_sequence.UNK_BASE * pad_l) + sequence + \
                (self.reference_sequence.UNK_BASE * pad_r)
        elif n > self.sequence_length:
            start = int((n - self.sequence_length) // 2)
            sequence = sequence[start:int(start + self.sequence_length)]
        sequence = str.upper(sequence)
        want = self._want(save_data)
        # scores stay in HBM: the writers format them there (only the mutation list comes to the host)
        import torch
        n_bases = mutate_n_bases if _honour_n_bases else 1
        muts, out, base = self.ism_scores(sequence, want, start_position, end_position, to_host=False,
                                          mutate_n_bases=n_bases)
        if n_bases == 1:
            pos, alt = muts[0].cpu().numpy(), muts[1].cpu().numpy()
            ids = [(str(int(p)), sequence[int(p)], "ACGT"[int(a)]) for p, a in zip(pos, alt)]
        else:
            from ._in_silico_mutagenesis import _ism_sample_id
            ids = [_ism_sample_id(sequence, mm) for mm in muts]
        results = []
        for s in want:
            if s == "predictions":
                arr = out["alt_predictions"]
                if output_format == "tsv":
                    # model_predict.py:789-792: the baseline row leads the predictions TSV
                    arr = torch.cat([base, arr], dim=0)
                    _out_dir, _p, sp = _scores_filepath(output_path_prefix, "predictions")
                    write_tsv(sp + ".tsv", ISM_COLS, self.features,
                              [("input_sequence", "NA", "NA")] + ids, arr)
                    continue
                results.append(("predictions", None, arr))
                _out_dir, _p, sp = _scores_filepath("{0}_ref".format(output_path_prefix), "predictions")
                write_hdf5(sp + ".h5", base)
                write_row_labels(os.path.join(_out_dir, "{0}_row_labels.txt".format(_p)), ["name"],
                                 [("input_sequence",)])
            else:
                results.append((s, None, out[s]))
        return _write_outputs(results, ids, self.features, ISM_COLS, output_path_prefix, output_format,
                              defer=_defer_writes) or None

    # ----------------------------------------------------------------------------------------------
    # predictions for sequences / FASTA / BED inputs (model_predict.py:265-597) and ISM over a FASTA file (:801-950)
    # ----------------------------------------------------------------------------------------------
    def _pad_or_truncate_sequence(self, sequence):
        """model_predict.py:1143-1153 with _pad_sequence / _truncate_sequence (_common.py:8-34)."""
        n = len(sequence)
        if n < self.sequence_length:
            diff = (self.sequence_length - n) / 2
            pad_l, pad_r = int(np.floor(diff)), math.ceil(diff)
            sequence = (self.reference_sequence.UNK_BASE * pad_l) + sequence + (self.reference_sequence.UNK_BASE * pad_r)
        elif n > self.sequence_length:
            start = int((n - self.sequence_length) // 2)
            sequence = sequence[start:int(start + self.sequence_length)]
        return sequence

    def _predict_codes(self, codes):
        """(n, L) uint8 codes on the device -> (n, F) CUDA predictions; rows are grouped into the reference's batches
        of ``batch_size`` (matters for DanQ's batch-axis recurrence)."""
        self.net.set_recurrence(self.batch_size, 1)
        return self.net.forward_codes(codes, precision=self.precision)

    def get_predictions_for_fasta_file(self, input_path, output_dir, output_format="tsv"):
        """model_predict.py:444-524: one prediction row per FASTA record, ids (index, name)."""
        import torch
        from ..sequences.genome import read_fasta
        os.makedirs(output_dir, exist_ok=True)
        _, filename = os.path.split(input_path)
        output_prefix = '.'.join(filename.split('.')[:-1])
        records = read_fasta(input_path)
        ids, rows = [], []
        for i, (name, arr) in enumerate(records):
            seq = self._pad_or_truncate_sequence(arr.tobytes().decode("latin-1"))
            rows.append(codes_of_string(seq))
            ids.append((i, name))
        codes = torch.from_numpy(np.stack(rows)).cuda() if rows else torch.zeros((0, self.sequence_length), dtype=torch.uint8).cuda()
        preds = self._predict_codes(codes) if len(rows) else torch.zeros((0, len(self.features)), device="cuda")
        _write_outputs([("predictions", None, preds)], ids, self.features, ["index", "name"],
                       os.path.join(output_dir, output_prefix), output_format)

    def _get_sequences_from_bed_file(self, input_path, strand_index=None, output_NAs_to_file=None,
                                     reference_sequence=None):
        """model_predict.py:265-345."""
        sequences, labels, na_rows = [], [], []
        check_chr = all(c.startswith("chr") for c in reference_sequence.get_chrs())
        chrs = set(reference_sequence.get_chrs())
        with open(input_path, 'r') as read_handle:
            for i, line in enumerate(read_handle):
                cols = line.strip().split('\t')
                if len(cols) < 3:
                    na_rows.append(line)
                    continue
                chrom, start, end = cols[0], cols[1], cols[2]
                strand = '.'
                if isinstance(strand_index, int) and len(cols) > strand_index:
                    strand = cols[strand_index]
                if 'chr' not in chrom and check_chr is True:
                    chrom = "chr{0}".format(chrom)
                if not str.isdigit(start) or not str.isdigit(end) or chrom not in chrs:
                    na_rows.append(line)
                    continue
                start, end = int(start), int(end)
                mid_pos = start + ((end - start) // 2)
                seq_start = mid_pos - self._start_radius
                seq_end = mid_pos + self._end_radius
                if reference_sequence:
                    if not reference_sequence.coords_in_bounds(chrom, seq_start, seq_end):
                        na_rows.append(line)
                        continue
                sequences.append((chrom, seq_start, seq_end, strand))
                labels.append((i, chrom, start, end, strand))
        if reference_sequence and output_NAs_to_file:
            with open(output_NAs_to_file, 'w') as file_handle:
                for na_row in na_rows:
                    file_handle.write(na_row)
        return sequences, labels

    def get_predictions_for_bed_file(self, input_path, output_dir, output_format="tsv", strand_index=None):
        """model_predict.py:347-442: predictions centred on every BED region, ids (index, chrom, start, end, strand,
        contains_unk); rows that cannot be fetched go to ``<prefix>.NA``."""
        import torch
        os.makedirs(output_dir, exist_ok=True)
        _, filename = os.path.split(input_path)
        output_prefix = '.'.join(filename.split('.')[:-1])
        G = self.reference_sequence
        seq_coords, labels = self._get_sequences_from_bed_file(
            input_path, strand_index=strand_index,
            output_NAs_to_file="{0}.NA".format(os.path.join(output_dir, output_prefix)), reference_sequence=G)
        n = len(labels)
        if n:
            strands = [c[3] for c in seq_coords]
            for st in strands:
                if st not in ('+', '-', '.'):
                    raise ValueError("Strand must be one of '+', '-', or '.'. Input was {0}".format(st))
            codes, flags = G.window_codes([c[0] for c in seq_coords], [c[1] for c in seq_coords], self.sequence_length,
                                          strands=strands)
            unk = ((flags & _lib.SB_FLAG_HAS_UPPER_N) != 0).cpu().numpy()
            preds = self._predict_codes(codes)
        else:
            unk = np.zeros(0, dtype=bool)
            preds = torch.zeros((0, len(self.features)), device="cuda")
        ids = [label + (bool(unk[i]),) for i, label in enumerate(labels)]
        for i, label in enumerate(labels):
            if unk[i]:
                warnings.warn(("For region {0}, reference sequence contains unknown base(s). --will be marked `True` "
                               "in the `contains_unk` column of the .tsv or row_labels .txt file.").format(label))
        _write_outputs([("predictions", None, preds)], ids, self.features,
                       ["index", "chrom", "start", "end", "strand", "contains_unk"],
                       os.path.join(output_dir, output_prefix), output_format)

    def get_predictions(self, input_path, output_dir=None, output_format="tsv", strand_index=None):
        """model_predict.py:526-597: a raw sequence (returns the (1, F) prediction), or a FASTA / BED file."""
        import torch
        if output_dir is None:
            sequence = self._pad_or_truncate_sequence(input_path)
            codes = torch.from_numpy(codes_of_string(sequence)[None]).cuda()
            return self._predict_codes(codes).cpu().numpy()
        elif input_path.endswith('.fa') or input_path.endswith('.fasta'):
            self.get_predictions_for_fasta_file(input_path, output_dir, output_format=output_format)
        else:
            self.get_predictions_for_bed_file(input_path, output_dir, output_format=output_format,
                                              strand_index=strand_index)
        return None

    def in_silico_mutagenesis_from_file(self, input_path, save_data, output_dir, mutate_n_bases=1,
                                        use_sequence_name=True, output_format="tsv", start_position=0,
                                        end_position=None):
        """model_predict.py:801-950: ISM of every record of a FASTA file, one set of output files per record
        (prefixed by the record name, or its index)."""
        from ..sequences.genome import read_fasta
        if end_position is None:
            end_position = self.sequence_length
        os.makedirs(output_dir, exist_ok=True)
        pending = []    # the files of record i are written while record i+1 is scored
        for i, (name, arr) in enumerate(read_fasta(input_path)):
            file_prefix = os.path.join(output_dir, name.replace(' ', '_') if use_sequence_name else str(i))
            nxt = self.in_silico_mutagenesis(arr.tobytes().decode("latin-1"), save_data, output_path_prefix=file_prefix,
                                             mutate_n_bases=mutate_n_bases, output_format=output_format,
                                             start_position=start_position, end_position=end_position,
                                             _defer_writes=True, _honour_n_bases=True)
            for finish in pending:
                finish()
            pending = nxt or []
        for finish in pending:
            finish()

    def variant_effect_prediction(self, vcf_file, save_data, output_dir=None, output_format="tsv",
                                  strand_index=None, require_strand=False):
        """model_predict.py:952-1141.  The VCF is parsed into a columnar ``VariantTable``; SNVs are scored straight from
        its arrays, other variants from host-built windows; the id columns of the output files are assembled from the
        table's field slices.  Under ``torch.distributed`` every rank scores a contiguous shard of the variants and
        (``self.multi_gpu_output == "shards"``, the default) writes its rows into the shared output files at its byte
        offset; ``"gather"`` collects the score tiles on rank 0 with one NCCL gather per score type instead."""
        import torch
        path, filename = os.path.split(vcf_file)
        output_path_prefix = '.'.join(filename.split('.')[:-1])
        if output_dir:
            os.makedirs(output_dir, exist_ok=True)
        else:
            output_dir = path
        output_path_prefix = os.path.join(output_dir, output_path_prefix)
        rank, world = 0, 1
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                rank, world = dist.get_rank(), dist.get_world_size()
        except ImportError:
            dist = None
        import time
        timing = self.last_timing = {}
        t_stage = time.perf_counter()

        def lap(name):
            nonlocal t_stage
            torch.cuda.synchronize()
            now = time.perf_counter()
            timing[name] = timing.get(name, 0.0) + now - t_stage
            t_stage = now
        G = self.reference_sequence
        table = VariantTable.from_vcf(vcf_file, strand_index=strand_index, require_strand=require_strand,
                                      output_NAs_to_file="{0}.NA".format(output_path_prefix) if rank == 0 else None,
                                      seq_context=(self._start_radius, self._end_radius), reference_sequence=G)
        lap("parse_s")
        want = self._want(save_data)
        n_all = len(table)
        # networks with a batch-axis recurrence (DanQ): the reference's batches are consecutive runs of
        # ``batch_size`` variants in file order, so shards start on batch boundaries and mixed SNV / indel files are
        # scored in one pass in file order
        recurrent = getattr(self.net, "has_recurrence", False)
        local, lo = table, 0
        if world > 1:
            from ..sharding import shard_range
            if recurrent:
                b_lo, b_hi = shard_range(-(-n_all // self.batch_size), rank, world)
                lo, hi = min(n_all, b_lo * self.batch_size), min(n_all, b_hi * self.batch_size)
            else:
                lo, hi = shard_range(n_all, rank, world)
            local = table.rows(lo, hi)
        n = len(local)
        F = len(self.features)
        names = [s for s in want if s != "predictions"]
        if "predictions" in want:
            names += ["ref_predictions", "alt_predictions"]
        dev = torch.device("cuda", int(self.net.device))
        out_dir, prefix = os.path.split(output_path_prefix)

        def output_files():
            """(score name, path without extension) in the reference's handler order."""
            files = []
            for s in want:
                if s == "predictions":
                    # write_ref_alt_handler.py:83-108: <prefix>.ref_predictions.* and <prefix>.alt_predictions.*
                    for tag in ("ref", "alt"):
                        p = os.path.join(out_dir, "{0}.{1}".format(prefix, tag) if prefix else tag)
                        files.append((tag + "_predictions", _scores_filepath(p, "predictions")[2], p))
                else:
                    files.append((s, _scores_filepath(output_path_prefix, s)[2], None))
            return files

        def enqueue(tbl):
            """Enqueues the scoring of ``tbl`` on the current stream.  Returns (dict of CUDA score arrays, resolve) where
            ``resolve()`` brings ref_match / contains_unk to the host (the only point that waits for the block)."""
            m = len(tbl)
            arrays = {k: torch.zeros((m, F), dtype=torch.float32, device=dev) for k in names}   # scores stay in HBM
            match = np.ones(m, dtype=bool)
            unk = np.zeros(m, dtype=bool)
            snv = tbl.is_snv.copy()
            if recurrent and not snv.all():
                snv[:] = False
            idx = np.flatnonzero(snv)
            r_snv = None
            if len(idx):
                G._unpicklable_init()
                to_genome = np.array([G._chrom_index[c] for c in tbl.chrom_names], dtype=np.int32)
                r_snv = self.score_snvs(to_genome[tbl.chrom_id[idx]], tbl.pos[idx], tbl.ref_code[idx], tbl.alt_code[idx],
                                        strands=tbl.minus[idx], save_data=want, to_host=False)
                if len(idx) == m:
                    for k in names:
                        arrays[k] = r_snv[k]
                else:
                    didx = torch.from_numpy(idx).to(dev)
                    for k in names:
                        arrays[k][didx] = r_snv[k]
            other = np.flatnonzero(~snv)
            if len(other):
                refs, alts = [], []
                for i in other:
                    rs, als, mt, u = build_ref_alt_strings(tbl.variant(i), G, self.sequence_length,
                                                           self._start_radius, self._end_radius)
                    refs.append(codes_of_string(rs)); alts.append(codes_of_string(als))
                    match[i], unk[i] = mt, u
                r = self.score_code_pairs(np.stack(refs), np.stack(alts), save_data=want, to_host=False)
                didx = torch.from_numpy(other).to(dev)
                for k in names:
                    arrays[k][didx] = r[k]

            def resolve():
                if r_snv is not None:
                    match[idx], unk[idx] = r_snv["ref_match"].cpu().numpy(), r_snv["contains_unk"].cpu().numpy()
                return match, unk
            return arrays, resolve

        def warn(tbl, match, unk):
            for i in np.flatnonzero(unk):
                warnings.warn("For variant ({0}, {1}, {2}, {3}, {4}, {5}), reference sequence contains "
                              "unknown base(s)--will be marked `True` in the `contains_unk` column of the "
                              ".tsv or the row_labels .txt file.".format(*tbl.variant(i)))
            for i in np.flatnonzero(~match):
                warnings.warn("For variant ({0}, {1}, {2}, {3}, {4}, {5}), reference does not match the "
                              "reference genome. Predictions/scores associated with this variant--where we "
                              "use '{3}' in the input sequence--will be marked `False` in the `ref_match` "
                              "column of the .tsv or the row_labels .txt file".format(*tbl.variant(i)))

        if world == 1 and output_format == "tsv":
            # ---- single process: blocks of variants are scored while the previous block's rows are formatted, read
            # back and written (the files grow block by block; byte-identical to writing everything at the end)
            from ._writers import TsvFile
            files = [(name, TsvFile(path + ".tsv", VARIANTEFFECT_COLS, self.features)) for name, path, _p in output_files()]
            block = max(self.batch_size, (self.vep_block_variants // self.batch_size) * self.batch_size)

            def host_lap(name, t0):
                now = time.perf_counter()
                timing[name] = timing.get(name, 0.0) + now - t0
                return now

            def finalize(tbl, arrays, resolve):
                t0 = time.perf_counter()
                match, unk = resolve()
                t0 = host_lap("host_wait_scores_s", t0)
                warn(tbl, match, unk)
                pre = tbl.id_prefix(match, unk)
                t0 = host_lap("host_id_columns_s", t0)
                for name, fh in files:
                    fh.append(arrays[name], pre[0], pre[1])
                host_lap("host_append_s", t0)
            pending = None
            for b0 in range(0, n, block):
                t0 = time.perf_counter()
                tbl = local.rows(b0, min(n, b0 + block))
                cur = (tbl,) + enqueue(tbl)
                host_lap("host_enqueue_s", t0)
                if pending is not None:
                    finalize(*pending)
                pending = cur
            if pending is not None:
                finalize(*pending)
            t0 = time.perf_counter()
            for _name, fh in files:
                fh.finish()
            host_lap("host_wait_writers_s", t0)
            lap("score_and_write_s")
            return None

        arrays, resolve = enqueue(local)
        match, unk = resolve()
        lap("score_s")
        warn(local, match, unk)
        lap("warnings_s")
        gather = world > 1 and (self.multi_gpu_output == "gather" or output_format != "tsv")
        if gather:
            from ..sharding import gather_tiles
            if recurrent:   # batch-aligned shards are not the equal split gather_tiles assumes
                raise NotImplementedError("multi_gpu_output='gather' is not available for recurrent networks; use 'shards'")
            for k in list(arrays.keys()):
                arrays[k] = gather_tiles(arrays[k], n_all, dst=0)
            mu = torch.from_numpy(np.stack([match, unk], axis=1).astype(np.float32)).to(dev)
            g = gather_tiles(mu, n_all, dst=0)
            if rank != 0:
                return None
            match, unk = g[:, 0].cpu().numpy() > 0.5, g[:, 1].cpu().numpy() > 0.5
            local = table
        results = [("predictions" if p is not None else name, p, arrays[name]) for name, _path, p in output_files()]
        lap("gather_s")
        if output_format == "tsv":
            pre = local.id_prefix(match, unk)
            lap("id_columns_s")
            finishers = [submit_tsv(_scores_filepath(output_path_prefix if p is None else p, name)[2] + ".tsv",
                                    VARIANTEFFECT_COLS, self.features, arr, pre[0], pre[1],
                                    dist_group=None if gather else "auto") for name, p, arr in results]
            lap("submit_s")
            for finish in finishers:
                finish()
            lap("write_s")
            if world > 1 and not gather:
                dist.barrier()      # every rank's rows are in the files when any rank returns
        else:
            ids = [local.variant(i) + (bool(match[i]), bool(unk[i])) for i in range(len(local))]
            _write_outputs(results, ids, self.features, VARIANTEFFECT_COLS, output_path_prefix, output_format)
        return None
