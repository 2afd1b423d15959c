"""Host side of kernels (c)+(d): turn a Selene model (``torch.nn.Module`` such as
``models/deepsea.py:DeepSEA`` of the reference) into an ``sb_net`` handle and run it.

``compile_module`` walks the module's children in definition order — the reference's models are
``nn.Sequential`` stacks of Conv1d / ReLU / MaxPool1d / Dropout / BatchNorm1d / Linear / Sigmoid
(models/deepsea.py:21-50, models/heartenn.py:21-67) — and emits the layer list of
``include/selene_b200.h``.  Eval-mode semantics only: Dropout is identity, BatchNorm1d uses running
statistics and is folded into the FOLLOWING conv/linear layer exactly (no padding anywhere).
"""
import ctypes as C

import numpy as np

from . import _lib


def layer(kind, weight, bias, relu=0, pool=1):
    return dict(type=kind, weight=np.ascontiguousarray(weight, dtype=np.float32),
                bias=np.ascontiguousarray(bias, dtype=np.float32), relu=int(relu), pool=int(pool))


def deepsea_layers_from_state(state, pools=(4, 4, 1)):
    """Layer list from a DeepSEA ``state_dict`` (keys conv_net.{0,4,8}, classifier.{0,2};
    models/deepsea.py:21-50)."""
    g = lambda k: state[k].detach().cpu().numpy()
    out = []
    for idx, pool in zip((0, 4, 8), pools):
        out.append(layer("conv", g("conv_net.%d.weight" % idx), g("conv_net.%d.bias" % idx), 1, pool))
    out.append(layer("fc", g("classifier.0.weight"), g("classifier.0.bias"), 1, 1))
    out.append(layer("fc", g("classifier.2.weight"), g("classifier.2.bias"), 0, 1))
    return out


def unwrap_non_strand_specific(model):
    """(inner model, strand mode) of a ``NonStrandSpecific`` wrapper (utils/non_strand_specific_module.py:26-85,
    recognised by its ``model`` / ``mode`` attributes), else (model, 0).  Mode 1 = mean, 2 = max."""
    if type(model).__name__ == "NonStrandSpecific" and hasattr(model, "model") and hasattr(model, "mode"):
        if model.mode not in ("mean", "max"):
            raise ValueError("Mode should be one of 'mean' or 'max' but was{0}.".format(model.mode))
        return model.model, 1 if model.mode == "mean" else 2
    return model, 0


def layers_from_module(model):
    """Generic walk over Sequential conv stacks (Conv1d / Conv2d with (1,k) kernels / Linear / ReLU /
    MaxPool / Dropout / BatchNorm1d / Sigmoid).  The final Sigmoid is implied by sb_net.  A NonStrandSpecific
    wrapper is looked through (pass ``strand_mode`` from ``unwrap_non_strand_specific`` to ``B200Net``)."""
    import torch.nn as nn
    model, _ = unwrap_non_strand_specific(model)
    mods = []

    def walk(m):
        kids = list(m.children())
        if not kids:
            mods.append(m)
        for k in kids:
            walk(k)
    walk(model)
    layers = []
    pending_bn = None  # (scale, shift) per channel to fold into the next conv/linear
    flat_channels = None
    for m in mods:
        if isinstance(m, (nn.Conv1d, nn.Conv2d)):
            w = m.weight.detach().cpu().numpy().astype(np.float64)
            if w.ndim == 4:  # Conv2d (1, k): models/seqweaver.py:30-38
                if w.shape[2] != 1:
                    raise ValueError("only (1,k) Conv2d kernels are supported")
                w = w[:, :, 0, :]
            b = (m.bias.detach().cpu().numpy().astype(np.float64) if m.bias is not None
                 else np.zeros(w.shape[0]))
            if pending_bn is not None:
                scale, shift = pending_bn
                b = b + (w * shift[None, :, None]).sum(axis=(1, 2))
                w = w * scale[None, :, None]
                pending_bn = None
            layers.append(layer("conv", w, b, 0, 1))
            flat_channels = w.shape[0]
        elif isinstance(m, nn.Linear):
            w = m.weight.detach().cpu().numpy().astype(np.float64)
            b = (m.bias.detach().cpu().numpy().astype(np.float64) if m.bias is not None
                 else np.zeros(w.shape[0]))
            if pending_bn is not None:
                scale, shift = pending_bn
                if scale.shape[0] != w.shape[1]:
                    # BatchNorm over conv channels followed by flatten: expand channel-major
                    rep = w.shape[1] // scale.shape[0]
                    scale, shift = np.repeat(scale, rep), np.repeat(shift, rep)
                b = b + w @ shift
                w = w * scale[None, :]
                pending_bn = None
            layers.append(layer("fc", w, b, 0, 1))
        elif isinstance(m, nn.LSTM):
            # models/danQ.py:29-31.  The reference feeds it (positions, batch, channels) with
            # batch_first=True, so the recurrence runs over the ROWS OF THE BATCH;
            # sb_net reproduces exactly that (sb_net_set_recurrence / B200Net.set_recurrence).
            if not (m.bidirectional and m.num_layers == 1 and m.bias):
                raise ValueError("only single-layer bidirectional LSTMs with bias are supported")
            if pending_bn is not None:
                raise ValueError("BatchNorm before an LSTM cannot be folded")
            g = lambda n: getattr(m, n).detach().cpu().numpy().astype(np.float32)
            layers.append(dict(type="bilstm", weight=g("weight_ih_l0"), bias=g("bias_ih_l0"), relu=0, pool=1,
                               aux=[g("weight_hh_l0"), g("bias_hh_l0"), g("weight_ih_l0_reverse"),
                                    g("bias_ih_l0_reverse"), g("weight_hh_l0_reverse"), g("bias_hh_l0_reverse")]))
        elif isinstance(m, nn.ReLU):
            if pending_bn is not None:
                raise ValueError("BatchNorm followed by ReLU cannot be folded forward")
            layers[-1]["relu"] = 1
        elif isinstance(m, (nn.MaxPool1d, nn.MaxPool2d)):
            k = m.kernel_size if isinstance(m.kernel_size, int) else m.kernel_size[-1]
            s = m.stride if isinstance(m.stride, int) else m.stride[-1]
            if k != s:
                raise ValueError("only MaxPool with stride == kernel is supported")
            if pending_bn is not None:
                raise ValueError("BatchNorm followed by MaxPool cannot be folded forward")
            layers[-1]["pool"] = int(k)
        elif isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
            g = m.weight.detach().cpu().numpy().astype(np.float64) if m.affine else 1.0
            be = m.bias.detach().cpu().numpy().astype(np.float64) if m.affine else 0.0
            mu = m.running_mean.detach().cpu().numpy().astype(np.float64)
            var = m.running_var.detach().cpu().numpy().astype(np.float64)
            scale = g / np.sqrt(var + m.eps)
            shift = be - mu * scale
            if pending_bn is not None:
                ps, pf = pending_bn
                scale, shift = scale * ps, scale * pf + shift
            pending_bn = (scale, shift)
        elif isinstance(m, (nn.Dropout, nn.Sigmoid, nn.Identity, nn.Flatten)):
            continue
        elif m.__class__.__name__ in ("Lambda", "LambdaBase"):
            continue  # reshape helpers of models/seqweaver.py:9-23,44-55
        else:
            raise ValueError("unsupported module %s" % type(m).__name__)
    if pending_bn is not None:
        raise ValueError("trailing BatchNorm cannot be folded")
    return layers


class B200Net(object):
    """An ``sb_net`` handle plus its reusable device workspace."""

    def __init__(self, layers, sequence_length, perm=(0, 1, 2, 3), device=None, strand_mode=0):
        import torch
        self._lib = _lib.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.sequence_length = int(sequence_length)
        self.perm = tuple(int(p) for p in perm)
        arr = (_lib.sb_layer * len(layers))()
        keep = []
        for i, l in enumerate(layers):
            w = np.ascontiguousarray(l["weight"], dtype=np.float32)
            b = np.ascontiguousarray(l["bias"], dtype=np.float32)
            keep += [w, b]
            arr[i].type = {"conv": _lib.SB_LAYER_CONV, "fc": _lib.SB_LAYER_FC,
                           "bilstm": _lib.SB_LAYER_BILSTM}[l["type"]]
            if l["type"] == "conv":
                arr[i].cout, arr[i].cin, arr[i].k = w.shape[0], w.shape[1], w.shape[2]
            elif l["type"] == "bilstm":
                arr[i].cout, arr[i].cin, arr[i].k = w.shape[0] // 4, w.shape[1], 1
                for a, t in enumerate(l["aux"]):
                    t = np.ascontiguousarray(t, dtype=np.float32)
                    keep.append(t)
                    arr[i].aux[a] = t.ctypes.data
            else:
                arr[i].cout, arr[i].cin, arr[i].k = w.shape[0], w.shape[1], 1
            arr[i].relu = int(l.get("relu", 0))
            arr[i].pool = int(l.get("pool", 1))
            arr[i].weight = w.ctypes.data
            arr[i].bias = b.ctypes.data
        h = C.c_void_p()
        _lib.check(self._lib.sb_net_create(arr, len(layers), self.sequence_length,
                                           _lib.perm_arg(self.perm), self.device, C.byref(h)),
                   "sb_net_create")
        self._h = h
        self.has_recurrence = any(l["type"] == "bilstm" for l in layers)   # outputs depend on row grouping (DanQ)
        self.n_outputs = int(self._lib.sb_net_n_outputs(h))
        self.flops_per_seq = float(self._lib.sb_net_flops_per_seq(h))
        self._ws = None
        self.strand_mode = 0
        if strand_mode:
            self.set_strand_mode(strand_mode)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.sb_net_destroy(h)
            except Exception:
                pass
            self._h = None

    def set_recurrence(self, group, interleave=1):
        """Row grouping of BiLSTM recurrences (DanQ): ``group`` = the caller's batch_size, ``interleave``
        = 2 when rows are interleaved ref/alt pairs.  No effect on networks without an LSTM."""
        _lib.check(self._lib.sb_net_set_recurrence(self._h, int(group), int(interleave)), "sb_net_set_recurrence")
        # the workspace requirement depends on the grouping; _workspace() re-queries it and only ever grows

    def set_strand_mode(self, mode):
        """0: plain; 1 / "mean", 2 / "max": the reference's NonStrandSpecific wrapper around the network."""
        mode = {"mean": 1, "max": 2}.get(mode, mode)
        _lib.check(self._lib.sb_net_set_strand_mode(self._h, int(mode)), "sb_net_set_strand_mode")
        self.strand_mode = int(mode)

    def _workspace(self, n, precision):
        import torch
        need = int(self._lib.sb_net_workspace_bytes(self._h, int(n), precision))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty((need,), dtype=torch.uint8, device=torch.device("cuda", self.device))
        return self._ws

    @staticmethod
    def _prec(precision):
        return {"fp32": _lib.SB_PREC_FP32, "bf16": _lib.SB_PREC_BF16}[precision]

    def forward_codes(self, codes, src=None, sub_pos=None, sub_code=None, n=None, precision="bf16",
                      out=None):
        """codes: uint8 CUDA (n_src, L).  Returns fp32 CUDA (n, F) sigmoid outputs."""
        import torch
        n = int(codes.shape[0] if n is None else n)
        p = self._prec(precision)
        ws = self._workspace(n, p)
        if out is None:
            out = torch.empty((n, self.n_outputs), dtype=torch.float32, device=codes.device)
        _lib.check(self._lib.sb_net_forward_codes(
            self._h, _lib.ptr(codes), _lib.ptr(src), _lib.ptr(sub_pos), _lib.ptr(sub_code), n, p,
            _lib.ptr(ws), ws.numel(), _lib.ptr(out), _lib.stream_ptr()), "sb_net_forward_codes")
        return out

    def forward_onehot(self, x, precision="bf16"):
        """x: fp32 CUDA (n, L, 4) (any values).  Returns fp32 CUDA (n, F)."""
        import torch
        x = x.contiguous()
        n = int(x.shape[0])
        p = self._prec(precision)
        ws = self._workspace(n, p)
        out = torch.empty((n, self.n_outputs), dtype=torch.float32, device=x.device)
        _lib.check(self._lib.sb_net_forward_onehot(self._h, _lib.ptr(x), n, p, _lib.ptr(ws), ws.numel(),
                                                   _lib.ptr(out), _lib.stream_ptr()),
                   "sb_net_forward_onehot")
        return out

    def forward_score(self, codes, src, sub_pos, sub_code, n, pair_mode, precision="bf16",
                      base_pred=None, base_index=None, want=("abs_diffs", "logits"), outs=None):
        """Fused forward + scoring.  Returns dict name -> fp32 CUDA tensor with names among
        abs_diffs, diffs, logits, ref_predictions, alt_predictions."""
        import torch
        n = int(n)
        p = self._prec(precision)
        ws = self._workspace(n, p)
        n_out = n // 2 if pair_mode == _lib.SB_PAIR_INTERLEAVED else n
        res = {}
        for name in want:
            if outs is not None and name in outs:
                res[name] = outs[name]
            else:
                res[name] = torch.empty((n_out, self.n_outputs), dtype=torch.float32, device=codes.device)
        g = lambda k: _lib.ptr(res[k]) if k in res else None
        _lib.check(self._lib.sb_net_forward_score(
            self._h, _lib.ptr(codes), _lib.ptr(src), _lib.ptr(sub_pos), _lib.ptr(sub_code), n, p,
            int(pair_mode), _lib.ptr(base_pred), _lib.ptr(base_index), _lib.ptr(ws), ws.numel(),
            g("abs_diffs"), g("diffs"), g("logits"), g("ref_predictions"), g("alt_predictions"),
            _lib.stream_ptr()), "sb_net_forward_score")
        return res
