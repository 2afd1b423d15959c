"""Packed training samples: the oracle's restatement of ``unpackbits_sequence`` / ``unpackbits_targets``
against the reference's own two functions (taken from the unmodified install in baseline/_ref — the module itself
imports h5py, which this image lacks, so the two function definitions are compiled on their own), and the CUDA
unpack / pack kernels + ``H5Dataset`` / ``H5DataLoader`` against the oracle."""
import ast
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import selene_oracle as O  # noqa: E402

REF_DL = os.path.join(ROOT, "baseline", "_ref", "selene_sdk", "samplers", "dataloader.py")


def _samples(n, L, F, seed, p_unk=0.03):
    """(n, L, 4) float one-hot with unknown rows = 0.25 (what Genome.get_encoding_from_coords returns), (n, F) 0/1."""
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, 4, (n, L))
    x = np.zeros((n, L, 4), dtype=np.float32)
    np.put_along_axis(x, codes[..., None], 1.0, axis=2)
    x[rng.random((n, L)) < p_unk] = 0.25
    y = (rng.random((n, F)) < 0.07).astype(np.float32)
    return x, y


def _reference_functions():
    if not os.path.exists(REF_DL):
        return None
    tree = ast.parse(open(REF_DL).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("unpackbits_sequence", "unpackbits_targets")]
    ns = {"np": np}
    exec(compile(ast.Module(body=keep, type_ignores=[]), REF_DL, "exec"), ns)
    return ns["unpackbits_sequence"], ns["unpackbits_targets"]


@pytest.mark.skipif(_reference_functions() is None, reason="baseline/_ref not installed")
@pytest.mark.parametrize("L,F", [(1000, 919), (37, 5), (8, 8), (1, 1)])
def test_oracle_unpack_equals_reference_functions(L, F):
    ref_seq, ref_tgt = _reference_functions()
    x, y = _samples(9, L, F, seed=L + F)
    ps, pt = O.pack_samples(x, y)
    assert ps.shape == (9, (L + 7) // 8, 4) and pt.shape == (9, (F + 7) // 8)
    assert np.array_equal(O.unpackbits_sequence(ps, L), ref_seq(ps, L))
    assert np.array_equal(O.unpackbits_sequence(ps[0], L), ref_seq(ps[0], L))
    assert np.array_equal(O.unpackbits_targets(pt, F), ref_tgt(pt, F))
    # the packed form is lossless for what the samplers produce
    assert np.array_equal(O.unpackbits_sequence(ps, L), x.astype(float))
    assert np.array_equal(O.unpackbits_targets(pt, F), y.astype(float))


@pytest.fixture(scope="module")
def torch():
    import torch as t
    return t


@pytest.mark.gpu
@pytest.mark.parametrize("L,F", [(1000, 919), (37, 5), (8, 8), (1, 1), (2003, 2002)])
def test_unpack_kernels_bit_exact(torch, L, F):
    from selene_b200.samplers import dataloader as D
    x, y = _samples(33, L, F, seed=3 * L + F)
    ps, pt = O.pack_samples(x, y)
    want_s, want_t = O.unpackbits_sequence(ps, L), O.unpackbits_targets(pt, F)
    got = D.unpack_sequences_device(ps, L).cpu().numpy()
    assert got.dtype == np.float32 and np.array_equal(got, want_s)
    got16 = D.unpack_sequences_device(ps, L, dtype="bfloat16").float().cpu().numpy()
    assert np.array_equal(got16, want_s)                                  # 0, 0.25 and 1 are exact in bf16
    assert np.array_equal(D.unpack_targets_device(pt, F).cpu().numpy(), want_t)
    got8 = D.unpack_targets_device(pt, F, dtype="uint8").cpu().numpy()
    assert got8.dtype == np.uint8 and np.array_equal(got8, want_t)
    # reference signatures (numpy in, float64 numpy out), single sample and batch
    assert np.array_equal(D.unpackbits_sequence(ps[1], L), want_s[1])
    assert np.array_equal(D.unpackbits_sequence(ps, L), want_s)
    assert np.array_equal(D.unpackbits_targets(pt, F), want_t)
    assert D.unpackbits_sequence(ps, L).dtype == np.float64
    # centre crop
    if L >= 8:
        a, n_out = L // 4, L // 2
        assert np.array_equal(D.unpack_sequences_device(ps, L, a, n_out).cpu().numpy(), want_s[:, a:a + n_out])
    with pytest.raises(ValueError):
        D.unpack_sequences_device(ps, L, 1, L)


@pytest.mark.gpu
@pytest.mark.parametrize("L,F", [(1000, 919), (37, 5), (8, 8), (1, 1), (250, 2002)])
def test_pack_kernels_equal_packbits(torch, L, F):
    from selene_b200.samplers import dataloader as D
    x, y = _samples(21, L, F, seed=L * 7 + F)
    ps, pt = O.pack_samples(x, y)
    gs, gt = D.pack_samples_device(x, y)
    assert np.array_equal(gs.cpu().numpy(), ps) and np.array_equal(gt.cpu().numpy(), pt)
    gs2, none = D.pack_samples_device(sequences=x)
    assert none is None and np.array_equal(gs2.cpu().numpy(), ps)
    none, gt2 = D.pack_samples_device(targets=y)
    assert none is None and np.array_equal(gt2.cpu().numpy(), pt)


@pytest.mark.gpu
def test_empty_inputs(torch):
    from selene_b200.samplers import dataloader as D
    assert tuple(D.unpack_sequences_device(np.zeros((0, 125, 4), np.uint8), 1000).shape) == (0, 1000, 4)
    assert tuple(D.unpack_targets_device(np.zeros((0, 115), np.uint8), 919).shape) == (0, 919)


@pytest.mark.gpu
def test_h5_dataset_and_loader_on_a_mapping(torch):
    """_H5Dataset.__getitem__ semantics (dataloader.py:226-265) incl. use_seq_len / shift, and the loader's batches
    covering every sample exactly once per epoch."""
    from selene_b200.samplers import H5DataLoader, H5Dataset
    L, F, n = 1000, 919, 300
    x, y = _samples(n, L, F, seed=11)
    ps, pt = O.pack_samples(x, y)
    db = {"sequences": ps, "targets": pt, "sequences_length": np.int64(L), "targets_length": np.int64(F)}
    ds = H5Dataset(db, unpackbits=True)
    assert len(ds) == n
    s, t = ds[17]
    assert s.dtype == torch.float32 and tuple(s.shape) == (L, 4) and np.array_equal(s.numpy(), x[17])
    assert np.array_equal(t.numpy(), y[17].astype(float))
    # centre window, shifted: mid - ceil(use/2) + shift ... mid + use//2 + shift
    ds2 = H5Dataset(db, unpackbits=True, use_seq_len=201, shift=-7)
    lo, hi = 500 - 101 - 7, 500 + 100 - 7
    s2, _ = ds2[5]
    assert np.array_equal(s2.numpy(), x[5, lo:hi])
    # unpacked storage is passed through
    ds3 = H5Dataset({"sequences": x, "targets": y.astype(np.uint8)})
    s3, t3 = ds3.batch([0, 299, 4])
    assert np.array_equal(s3.cpu().numpy(), x[[0, 299, 4]]) and np.array_equal(t3.cpu().numpy(), y[[0, 299, 4]])
    # loader: every sample once per epoch, tensors on the device, reshuffled between epochs
    dl = H5DataLoader(ds, batch_size=64, shuffle=True, seed=3)
    assert len(dl) == 5
    seen, first = [], None
    for bs, bt in dl:
        assert bs.is_cuda and bt.is_cuda and bs.shape[1:] == (L, 4) and bt.shape[1] == F
        if first is None:
            first = bs.clone()
        key = bt.cpu().numpy().astype(np.uint8)
        for row_s, row_t in zip(bs.cpu().numpy(), key):
            hit = np.flatnonzero((y == row_t).all(axis=1))
            hit = [h for h in hit if np.array_equal(x[h], row_s)]
            assert len(hit) >= 1
            seen.append(hit[0])
    assert len(seen) == n and len(set(seen)) == n
    again = next(iter(dl))[0]
    assert not torch.equal(again, first)
    sub = H5DataLoader(ds, batch_size=50, use_subset=(100, 200))
    idx = []
    for bs, bt in sub:
        for row_s in bs.cpu().numpy():
            idx.append(next(h for h in range(100, 200) if np.array_equal(x[h], row_s)))
    assert sorted(idx) == list(range(100, 200))
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="h5py"):
            len(H5Dataset("/nonexistent/file.h5", unpackbits=True))
