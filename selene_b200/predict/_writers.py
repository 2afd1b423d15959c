"""Output files of the scoring entry points — SURVEY §8 a18 / f1 (reference: predict_handlers/handler.py:15-96,
190-235: TSV rows ``id columns + "{:.2e}" per value``; HDF5 dataset ``data`` float64 (n, F) + ``row_labels.txt``).

TSV bodies never exist on the host as Python objects: the score matrix stays in HBM, ``sb_format_e2_row_bytes``
gives every row's byte count, a prefix sum turns them into FILE OFFSETS, and a pool of writer threads — each with
its own CUDA stream, device text buffer and pinned host buffer — formats a chunk of rows on the device
(``sb_format_e2_rows_prefixed``: id columns and numbers in one pass), reads it back and ``pwrite``s it at its offset.
Because every chunk's offset is known up front, chunks of one file are written concurrently, and under
``torch.distributed`` every rank writes ITS shard of the rows into the one output file at
``header + sum(bytes of lower ranks)``: the only collective of the multi-GPU file path is an all-gather of one
integer per file (no score tile crosses NVLink or funnels through rank 0's PCIe link).  The files are
byte-identical to a single-process run.
"""
import mmap
import os
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .. import _lib

_CHUNK_BYTES = 32 << 20
_POOL = None
_POOL_SIZE = int(os.environ.get("SB_WRITER_THREADS", "0")) or max(2, min(12, (os.cpu_count() or 4) - 2))
_TLS = threading.local()
_USE_MMAP = os.environ.get("SB_WRITER_MMAP", "0") == "1"     # chunk text copied through a shared mapping instead of pwrite


def writer_pool():
    """Persistent writer threads (their streams, device and pinned buffers are allocated once)."""
    global _POOL
    if _POOL is None:
        _POOL = ThreadPoolExecutor(max_workers=_POOL_SIZE, thread_name_prefix="sb_writer")
    return _POOL


def prefixes_from_ids(ids):
    """(uint8 buffer, offsets) of ``"\\t".join(str(i) for i in info) + "\\t"`` per row of an id list (handler.py:36-42)."""
    chunks = [("\t".join(str(i) for i in info) + "\t").encode() for info in ids]
    off = np.zeros(len(chunks) + 1, dtype=np.int64)
    np.cumsum(np.fromiter((len(c) for c in chunks), dtype=np.int64, count=len(chunks)), out=off[1:])
    return np.frombuffer(b"".join(chunks), dtype=np.uint8), off


def _thread_state(device_index, need):
    """This writer thread's stream, device text buffer and pinned buffer (grown on demand)."""
    import torch
    st = getattr(_TLS, "state", None)
    if st is None or st["device"] != device_index:
        torch.cuda.set_device(device_index)
        st = _TLS.state = dict(device=device_index, stream=torch.cuda.Stream(device=device_index), dev=None, host=None)
    if st["dev"] is None or st["dev"].numel() < need:
        size = max(need, _CHUNK_BYTES + (4 << 20))
        st["dev"] = torch.empty(size, dtype=torch.uint8, device=torch.device("cuda", device_index))
        st["host"] = torch.empty(size, dtype=torch.uint8, pin_memory=True)
    return st


def _chunk_job(fd, file_off, data, r0, r1, row_off_dev, off0, nbytes, pre_dev, pre_off_dev, ready):
    """Formats rows [r0, r1) on this thread's stream and writes their ``nbytes`` of text at ``file_off``."""
    import torch
    lib = _lib.load()
    dev = data.device.index or 0
    st = _thread_state(dev, nbytes)
    F = int(data.shape[1])
    with torch.cuda.stream(st["stream"]):
        st["stream"].wait_event(ready)                      # the scores were produced on another stream
        # the kernel writes row r at out + row_off[r]: hand it the chunk buffer shifted back by the chunk's first offset
        out_ptr = st["dev"].data_ptr() - off0
        _lib.check(lib.sb_format_e2_rows_prefixed(
            data.data_ptr() + r0 * F * 4, r1 - r0, F, row_off_dev.data_ptr() + r0 * 8, pre_dev.data_ptr(),
            pre_off_dev.data_ptr() + r0 * 8, out_ptr, st["stream"].cuda_stream), "sb_format_e2_rows_prefixed")
        st["host"][:nbytes].copy_(st["dev"][:nbytes], non_blocking=True)
    st["stream"].synchronize()
    if _USE_MMAP and nbytes:
        # the file already has its size (ftruncate): copy into a shared mapping of the chunk's pages.  Unlike pwrite,
        # page-cache copies through a mapping do not take the inode's write lock, so chunks of ONE file land in parallel.
        lo = file_off - file_off % mmap.ALLOCATIONGRANULARITY
        with mmap.mmap(fd, file_off + nbytes - lo, offset=lo) as m:
            dst = np.frombuffer(m, dtype=np.uint8, count=nbytes, offset=file_off - lo)
            np.copyto(dst, st["host"].numpy()[:nbytes])
            del dst
        return nbytes
    view = memoryview(st["host"].numpy())[:nbytes]
    done = 0
    while done < nbytes:
        done += os.pwrite(fd, view[done:], file_off + done)
    return nbytes


def submit_tsv(path, columns_for_ids, features, data, prefix, prefix_off, dist_group="auto"):
    """Starts writing ``path``: header ``columns_for_ids + features`` and one row per row of ``data`` (CUDA float32
    (n, F) tensor, or anything convertible).  ``prefix`` / ``prefix_off``: the id-column bytes of every row
    (``prefixes_from_ids`` or ``VariantTable.id_prefix``).  Returns a ``finish()`` callable that waits for the
    chunks, closes the file and re-raises a writer's exception.

    Under an initialised ``torch.distributed`` group (``dist_group="auto"``) the call is collective: every rank
    passes its own shard of rows (rank order = row order) and all of them write into the same file."""
    import torch
    lib = _lib.load()
    t = data if isinstance(data, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32))
    t = t.to(device="cuda", dtype=torch.float32).contiguous()
    n, F = int(t.shape[0]), int(t.shape[1])
    dev = t.device
    header = "{0}\n".format("\t".join(list(columns_for_ids) + list(features))).encode()
    prefix_off = np.ascontiguousarray(prefix_off, dtype=np.int64)
    assert len(prefix_off) == n + 1, "one id prefix per row is needed"
    # ---- byte count of every row -> offsets inside this rank's text
    row_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    pre_dev = torch.from_numpy(np.array(prefix, dtype=np.uint8)).to(dev) if len(prefix) else \
        torch.zeros(1, dtype=torch.uint8, device=dev)
    pre_off_dev = torch.from_numpy(prefix_off).to(dev)
    if n:
        sizes = torch.empty(n, dtype=torch.int64, device=dev)
        _lib.check(lib.sb_format_e2_row_bytes(_lib.ptr(t), n, F, _lib.ptr(sizes), _lib.stream_ptr()),
                   "sb_format_e2_row_bytes")
        sizes += pre_off_dev[1:] - pre_off_dev[:-1]
        torch.cumsum(sizes, 0, out=row_off[1:])
    ready = torch.cuda.Event()
    ready.record()                                           # writer streams wait for the scores and the offsets
    off_host = row_off.cpu().numpy()
    total = int(off_host[-1])
    # ---- where this rank's text starts in the file
    rank, world, base, file_size = 0, 1, len(header), len(header) + total
    dist = None
    if dist_group == "auto":
        import torch.distributed as d
        if d.is_available() and d.is_initialized() and d.get_world_size() > 1:
            dist = d
    if dist is not None:
        rank, world = dist.get_rank(), dist.get_world_size()
        totals = torch.zeros(world, dtype=torch.int64, device=dev)
        totals[rank] = total
        dist.all_reduce(totals)                              # the one collective of the file path: 8 bytes per rank
        totals = totals.cpu().numpy()
        base = len(header) + int(totals[:rank].sum())
        file_size = len(header) + int(totals.sum())
    if rank == 0:
        fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644)
        os.pwrite(fd, header, 0)
        os.ftruncate(fd, file_size)
    if dist is not None:
        dist.barrier()                                       # the file exists with its final size
    if rank != 0:
        fd = os.open(path, os.O_RDWR)
    # ---- chunks of <= 32 MB of text, written concurrently at their offsets
    futures = []
    r0 = 0
    pool = writer_pool()
    while r0 < n:
        r1 = int(np.searchsorted(off_host, off_host[r0] + _CHUNK_BYTES, side="right")) - 1
        r1 = min(n, max(r1, r0 + 1))
        nbytes = int(off_host[r1] - off_host[r0])
        futures.append(pool.submit(_chunk_job, fd, base + int(off_host[r0]), t, r0, r1, row_off, int(off_host[r0]),
                                   nbytes, pre_dev, pre_off_dev, ready))
        r0 = r1

    def finish():
        err = None
        for f in futures:
            try:
                f.result()
            except Exception as e:       # keep draining so the descriptor is closed exactly once
                err = err or e
        os.close(fd)
        if err is not None:
            raise err
        return path
    return finish


class TsvFile(object):
    """A TSV output that is appended to BLOCK BY BLOCK while later blocks are still being scored (single process):
    ``append`` sizes the block's rows on the device, places it after the text written so far and hands its chunks to the
    writer pool; ``finish`` waits for them.  The result is byte-identical to one ``submit_tsv`` over all rows."""

    def __init__(self, path, columns_for_ids, features):
        self.path = path
        header = "{0}\n".format("\t".join(list(columns_for_ids) + list(features))).encode()
        self._fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644)
        os.pwrite(self._fd, header, 0)
        self._pos = len(header)
        self._futures = []

    def append(self, data, prefix, prefix_off):
        import torch
        lib = _lib.load()
        t = data.to(dtype=torch.float32).contiguous()
        n, F = int(t.shape[0]), int(t.shape[1])
        if n == 0:
            return
        dev = t.device
        prefix_off = np.ascontiguousarray(prefix_off, dtype=np.int64)
        pre_dev = torch.from_numpy(np.array(prefix, dtype=np.uint8)).to(dev) if len(prefix) else \
            torch.zeros(1, dtype=torch.uint8, device=dev)
        pre_off_dev = torch.from_numpy(prefix_off).to(dev)
        sizes = torch.empty(n, dtype=torch.int64, device=dev)
        _lib.check(lib.sb_format_e2_row_bytes(_lib.ptr(t), n, F, _lib.ptr(sizes), _lib.stream_ptr()), "sb_format_e2_row_bytes")
        sizes += pre_off_dev[1:] - pre_off_dev[:-1]
        row_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        torch.cumsum(sizes, 0, out=row_off[1:])
        ready = torch.cuda.Event()
        ready.record()
        off_host = row_off.cpu().numpy()                       # waits for this block's scores only
        pool = writer_pool()
        r0 = 0
        while r0 < n:
            r1 = int(np.searchsorted(off_host, off_host[r0] + _CHUNK_BYTES, side="right")) - 1
            r1 = min(n, max(r1, r0 + 1))
            nbytes = int(off_host[r1] - off_host[r0])
            if _USE_MMAP and r0 == 0:
                os.ftruncate(self._fd, self._pos + int(off_host[-1]))
            self._futures.append(pool.submit(_chunk_job, self._fd, self._pos + int(off_host[r0]), t, r0, r1, row_off,
                                             int(off_host[r0]), nbytes, pre_dev, pre_off_dev, ready))
            r0 = r1
        self._pos += int(off_host[-1])

    def finish(self):
        err = None
        for f in self._futures:
            try:
                f.result()
            except Exception as e:
                err = err or e
        os.close(self._fd)
        if err is not None:
            raise err
        return self.path


def write_tsv(path, columns_for_ids, features, ids, data, prefix=None, prefix_off=None):
    """Blocking form of ``submit_tsv``; ``ids`` (list of tuples) is used when no prefix bytes are given."""
    if prefix is None:
        prefix, prefix_off = prefixes_from_ids(ids)
    return submit_tsv(path, columns_for_ids, features, data, prefix, prefix_off)()


def write_hdf5(path, data):
    """Dataset ``data`` float64 (n_rows, n_features) (handler.py:205-218).  The float64 staging copy is made on the
    device and read back through one transfer."""
    try:
        import h5py
    except ImportError as e:
        raise ImportError("output_format='hdf5' needs the h5py package, which is not installed in this environment; "
                          "use output_format='tsv'") from e
    if hasattr(data, "is_cuda"):
        data = data.double().cpu().numpy()
    with h5py.File(path, "w") as fh:
        fh.create_dataset("data", data=np.asarray(data, dtype=np.float64))


def write_row_labels(path, columns_for_ids, ids):
    with open(path, "w") as fh:
        fh.write("{0}\n".format("\t".join(columns_for_ids)))
        for info in ids:
            fh.write("{0}\n".format("\t".join(str(i) for i in info)))
