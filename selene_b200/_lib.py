"""ctypes binding of ``libselene_b200.so`` (the C ABI declared in ``include/selene_b200.h``).

This is the only module that touches the shared library.  There is deliberately no CPU fallback:
if the library is missing, or a call fails, the error is raised to the caller.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SELENE_B200_LIB points the binding at another build of the same C ABI (the mutation check of the parity suite,
# scripts/mutation_check.sh); there is still no fallback: a missing file raises.
LIB_PATH = os.environ.get("SELENE_B200_LIB") or os.path.join(_HERE, "libselene_b200.so")

SB_OK = 0
SB_F32, SB_BF16, SB_U8, SB_I64, SB_F64 = 0, 1, 2, 3, 4
SB_FLAG_HAS_UPPER_N, SB_FLAG_HAS_UNK = 1, 2
SB_LAYER_CONV, SB_LAYER_FC, SB_LAYER_BILSTM = 1, 2, 3
SB_PREC_FP32, SB_PREC_BF16 = 0, 1
SB_PAIR_INTERLEAVED, SB_PAIR_BASELINE = 0, 1


class SbError(RuntimeError):
    pass


class sb_layer(C.Structure):
    _fields_ = [("type", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32), ("k", C.c_int32),
                ("relu", C.c_int32), ("pool", C.c_int32),
                ("weight", C.c_void_p), ("bias", C.c_void_p), ("aux", C.c_void_p * 6)]


_P = C.c_void_p
_I = C.c_int
_L = C.c_int64

# name -> (restype, argtypes); mirrors include/selene_b200.h one to one
SIGNATURES = {
    "sb_version": (_I, []),
    "sb_last_error": (_I, [C.c_char_p, _I]),
    "sb_launch_count": (C.c_ulonglong, []),
    "sb_device_info": (_I, [_I, _P, _P, _P]),
    "sb_genome_create": (_I, [_P, _P, _P, _L, _P, _P, _I, _I, _P]),
    "sb_genome_destroy": (_I, [_P]),
    "sb_genome_codes": (_I, [_P, _P, _P, _P, _I, _I, _P, _P, _P]),
    "sb_encode_windows": (_I, [_P, _P, _P, _P, _I, _I, _P, _I, _P, _P, _P]),
    "sb_encode_codes": (_I, [_P, _L, _I, _P, _I, _P, _P]),
    "sb_ism_enumerate": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _I, _P]),
    "sb_ism_expand": (_I, [_P, _I, _P, _P, _P, _L, _P, _I, _P, _P]),
    "sb_snv_pairs": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P, _P]),
    "sb_intervals_create": (_I, [_P, _P, _P, _P, _L, _I, _I, _I, _P]),
    "sb_intervals_destroy": (_I, [_P]),
    "sb_label_bins": (_I, [_P, _P, _P, _P, _I, _P, _I, _P, _P]),
    "sb_net_create": (_I, [_P, _I, _I, _P, _I, _P]),
    "sb_net_destroy": (_I, [_P]),
    "sb_net_n_outputs": (_I, [_P]),
    "sb_net_flops_per_seq": (C.c_double, [_P]),
    "sb_net_set_recurrence": (_I, [_P, _I, _I]),
    "sb_net_set_strand_mode": (_I, [_P, _I]),
    "sb_net_set_timing": (_I, [_P, _I]),
    "sb_net_n_layers": (_I, [_P]),
    "sb_net_layer_flops_per_seq": (C.c_double, [_P, _I]),
    "sb_net_layer_time": (_I, [_P, _I, _P, _P]),
    "sb_net_workspace_bytes": (C.c_size_t, [_P, _L, _I]),
    "sb_net_forward_codes": (_I, [_P, _P, _P, _P, _P, _L, _I, _P, C.c_size_t, _P, _P]),
    "sb_net_forward_onehot": (_I, [_P, _P, _L, _I, _P, C.c_size_t, _P, _P]),
    "sb_net_forward_score": (_I, [_P, _P, _P, _P, _P, _L, _I, _I, _P, _P, _P, C.c_size_t,
                                  _P, _P, _P, _P, _P, _P]),
    "sb_net_train_init": (_I, [_P]),
    "sb_net_n_params": (C.c_size_t, [_P]),
    "sb_net_param_offset": (C.c_size_t, [_P, _I]),
    "sb_net_set_grad_event": (_I, [_P, _I, _P]),
    "sb_net_set_dropout": (_I, [_P, _I, C.c_float]),
    "sb_net_train_workspace_bytes": (C.c_size_t, [_P, _L]),
    "sb_net_forward_backward": (_I, [_P, _P, _P, _L, _P, C.c_ulonglong, _P, C.c_size_t, _P, _P, _P]),
    "sb_net_train_init_tc": (_I, [_P]),
    "sb_net_train_workspace_bytes_tc": (C.c_size_t, [_P, _L]),
    "sb_net_forward_backward_tc": (_I, [_P, _P, _P, _L, _P, C.c_ulonglong, _P, C.c_size_t, _P, _P, _P]),
    "sb_net_sgd_step": (_I, [_P, _P, C.c_float, C.c_float, C.c_float, _P]),
    "sb_net_sgd_step_layers": (_I, [_P, _P, _I, C.c_float, C.c_float, C.c_float, _I, _I, _I, _P]),
    "sb_net_get_layer_params": (_I, [_P, _I, _P, _P, _P]),
    "sb_net_get_layer_momentum": (_I, [_P, _I, _P, _P]),
    "sb_net_set_layer_momentum": (_I, [_P, _I, _P, _P]),
    "sb_score_pairs": (_I, [_P, _P, _P, _L, _L, _I, _P, _P, _P, _I, _P]),
    "sb_cast_f32_to_bf16": (_I, [_P, _L, _P, _P]),
    "sb_cast_bf16_to_f32": (_I, [_P, _L, _P, _P]),
    "sb_unpack_sequences": (_I, [_P, _L, _I, _I, _I, _I, _P, _P]),
    "sb_unpack_targets": (_I, [_P, _L, _I, _I, _P, _P]),
    "sb_pack_samples": (_I, [_P, _P, _L, _I, _I, _P, _P, _P]),
    "sb_format_e2_row_bytes": (_I, [_P, _L, _I, _P, _P]),
    "sb_format_e2_rows": (_I, [_P, _L, _I, _P, _P, _P]),
    "sb_format_e2_rows_prefixed": (_I, [_P, _L, _I, _P, _P, _P, _P, _P]),
    "sb_tc_wgrad_selftest": (_I, [_P, _P, _L, _I, _I, _I, _P, _P]),
    "sb_tc_gemm_selftest": (_I, [_P, _P, _I, _I, _I, _P, _P]),
}

_lib = None


def load():
    """Load the shared library (once) and declare every prototype.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SbError(
            "selene_b200: %s is missing — build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    buf = C.create_string_buffer(1024)
    load().sb_last_error(buf, 1024)
    return buf.value.decode("utf-8", "replace")


def check(rc, what):
    if rc != SB_OK:
        raise SbError("%s failed (code %d): %s" % (what, rc, last_error()))


def ptr(t):
    """Device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def perm_arg(perm):
    return (C.c_uint8 * 4)(*[int(x) for x in perm])
