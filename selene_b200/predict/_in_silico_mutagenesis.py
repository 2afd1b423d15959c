"""In-silico mutagenesis mutant generation on the GPU (reference:
selene_sdk/predict/_in_silico_mutagenesis.py).

``in_silico_mutagenesis_sequences`` and ``mutate_sequence`` keep the reference signatures; the
single-base case — the only one ``AnalyzeSequences.in_silico_mutagenesis`` ever uses
(model_predict.py:763-767 hard-wires ``mutate_n_bases=1``) — runs ``sb_ism_enumerate`` /
``sb_ism_expand``.  Multi-base combinations are an enumeration problem, not a data-parallel one,
and stay a host generator with the reference's ordering.
"""
import itertools

import numpy as np

from .. import _lib
from ..sequences import Genome, codes_of_string


def _check_range(sequence, mutate_n_bases, start_position, end_position):
    """Argument checks of _in_silico_mutagenesis.py:64-89 (same messages)."""
    if start_position >= end_position:
        raise ValueError(("Starting positions must be less than the ending "
                          "positions. Found a starting position of {0} with "
                          "an ending position of {1}.").format(start_position, end_position))
    if start_position < 0:
        raise ValueError("Negative starting positions are not supported.")
    if end_position < 0:
        raise ValueError("Negative ending positions are not supported.")
    if start_position >= len(sequence):
        raise ValueError(("Starting positions must be less than the sequence length."
                          " Found a starting position of {0} with a sequence length "
                          "of {1}.").format(start_position, len(sequence)))
    if end_position > len(sequence):
        raise ValueError(("Ending positions must be less than or equal to the sequence "
                          "length. Found an ending position of {0} with a sequence "
                          "length of {1}.").format(end_position, len(sequence)))
    if (end_position - start_position) < mutate_n_bases:
        raise ValueError(("Fewer bases exist in the substring specified by the starting "
                          "and ending positions than need to be mutated. There are only "
                          "{0} currently, but {1} bases must be mutated at a "
                          "time").format(end_position - start_position, mutate_n_bases))


def ism_mutations_device(sequence, start_position=0, end_position=None, reference_sequence=Genome):
    """Single-base mutations of ``sequence`` enumerated on the device.

    Returns ``(mut_pos int32 (M), mut_alt uint8 (M) canonical codes, codes uint8 (1, L))`` CUDA tensors
    in the reference's order (_in_silico_mutagenesis.py:91-107)."""
    import torch
    if end_position is None:
        end_position = len(sequence)
    _check_range(sequence, 1, start_position, end_position)
    # the reference compares characters with BASES_ARR (uppercase): a lowercase or unknown
    # reference character matches none of them and yields all four alternatives
    raw = np.frombuffer(sequence.encode("latin-1", "replace"), dtype=np.uint8)
    codes_np = codes_of_string(sequence).copy()
    codes_np[(raw >= 97) & (raw <= 122)] = 4
    L = len(sequence)
    enum_codes = torch.from_numpy(codes_np).cuda()
    cap = 4 * (end_position - start_position)
    pos = torch.empty((cap,), dtype=torch.int32, device="cuda")
    alt = torch.empty((cap,), dtype=torch.uint8, device="cuda")
    cnt = torch.zeros((1,), dtype=torch.int32, device="cuda")
    _lib.check(_lib.load().sb_ism_enumerate(_lib.ptr(enum_codes), L, int(start_position), int(end_position),
                                            _lib.perm_arg(reference_sequence.perm()), _lib.ptr(pos),
                                            _lib.ptr(alt), _lib.ptr(cnt), cap, _lib.stream_ptr()),
               "sb_ism_enumerate")
    m = int(cnt.item())
    true_codes = torch.from_numpy(codes_of_string(sequence).copy()).cuda().reshape(1, L)
    return pos[:m], alt[:m], true_codes


def expand_mutants_device(codes, mut_pos, mut_alt, src=None, reference_sequence=Genome, dtype="float32"):
    """``mutate_sequence`` for a whole batch: CUDA (M, L, 4) one-hot mutants."""
    import torch
    m = int(mut_pos.numel())
    L = int(codes.shape[-1])
    tdt = torch.float32 if dtype == "float32" else torch.bfloat16
    out = torch.empty((m, L, 4), dtype=tdt, device=codes.device)
    _lib.check(_lib.load().sb_ism_expand(_lib.ptr(codes), L, _lib.ptr(src), _lib.ptr(mut_pos), _lib.ptr(mut_alt),
                                         m, _lib.perm_arg(reference_sequence.perm()),
                                         _lib.SB_F32 if dtype == "float32" else _lib.SB_BF16, _lib.ptr(out),
                                         _lib.stream_ptr()), "sb_ism_expand")
    return out


def in_silico_mutagenesis_sequences(sequence, mutate_n_bases=1, reference_sequence=Genome,
                                    start_position=0, end_position=None):
    """Same contract as _in_silico_mutagenesis.py:8-107: list of ``[(pos, base), ...]``."""
    if end_position is None:
        end_position = len(sequence)
    _check_range(sequence, mutate_n_bases, start_position, end_position)
    if mutate_n_bases == 1:
        pos, alt, _ = ism_mutations_device(sequence, start_position, end_position, reference_sequence)
        pos, alt = pos.cpu().numpy(), alt.cpu().numpy()
        return [[(int(p), "ACGT"[int(a)])] for p, a in zip(pos, alt)]
    sequence_alts = [[b for b in reference_sequence.BASES_ARR if b != ref] for ref in sequence]
    out = []
    for indices in itertools.combinations(range(start_position, end_position), mutate_n_bases):
        for mutations in itertools.product(*[sequence_alts[i] for i in indices]):
            out.append(list(zip(indices, mutations)))
    return out


def mutate_sequence(encoding, mutation_information, reference_sequence=Genome):
    """_in_silico_mutagenesis.py:110-143 for one encoding (host arrays in, host array out; the batched
    device form is ``expand_mutants_device``)."""
    mutated_seq = np.copy(encoding)
    for (position, alt) in mutation_information:
        replace_base = reference_sequence.BASE_TO_INDEX[alt]
        mutated_seq[position, :] = 0
        mutated_seq[position, replace_base] = 1
    return mutated_seq


def _ism_sample_id(sequence, mutation_information):
    """_in_silico_mutagenesis.py:146-170."""
    positions, refs, alts = [], [], []
    for (position, alt) in mutation_information:
        positions.append(str(position))
        refs.append(sequence[position])
        alts.append(alt)
    return (';'.join(positions), ';'.join(refs), ';'.join(alts))
