"""bench.py — measures BASELINE.json's metric, "mutant+variant predictions/sec (DeepSEA 1kb, 919 tgt)",
on the workload of configs[1]: DeepSEA variant_effect_prediction over synthetic SNVs on a synthetic
genome, abs_diffs + logits scoring.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path (SNV windows -> ref/alt rows -> DeepSEA forward -> scores) over
one block of --variants-per-step variants (2 predictions each) per GPU.  The variant stream is generated
in blocks: block b of rank r is seeded by (2024, r, b), so both arms see the same variants.  One JSON line
is printed by rank 0.  Legs of our arm:

  value      device-resident inputs, weak scaling (every rank its own blocks, no data-path collective)
  e2e        the same through AnalyzeSequences.prepare_snvs + score_prepared from pinned HOST arrays, with the
             120 MB/step of score matrices copied back to pinned host memory, all inside the timed region
  roofline   dominant tcgen05 layer kernel: algorithmic FLOPs per launch / CUDA-event time per launch
  roofline_hbm  (N=1) the HBM-bound kernels of stage 1/2a timed alone against the measured copy bandwidth
  strong     configs[1] as BASELINE.json states it: a FIXED list of --strong-variants SNVs (default 2^20) split
             into contiguous shards over the ranks, timed (a) with per-rank device->host shard read-back and
             (b) with one NCCL gather of the score tiles to rank 0 + rank 0's device->host read
  train      configs[2]: the DeepSEA training step, batch 64 per GPU, gradients all-reduced over NCCL (weak scaling)
  cpu_baseline  (N=1) the unmodified reference (baseline/_ref) on a bounded prefix of the first timed block,
             on all host threads; the oracle port's rate on the same prefix is reported beside it

The ``--impl reference`` arm times the UNMODIFIED reference's own AnalyzeSequences.variant_effect_prediction
(baseline/_ref, torch CPU, all host threads) on the same genome and the first --ref-variants-per-step variants
of every block of the same stream.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mutant+variant predictions/sec (DeepSEA 1kb, 919 tgt)"
UNIT = "predictions/s"
SEQ_LEN, N_FEATURES = 1000, 919
STREAM_SEED = 2024


# --------------------------------------------------------------------------------------------------
# synthetic workload (scaled by --genome-mbp)
# --------------------------------------------------------------------------------------------------
def synth_genome_arrays(total_mbp, seed=1337):
    """24 contigs chr1..chr22,chrX,chrY of equal length; iid ACGT, 10 kb N telomeres, 0.5% N in runs
    of 50-5000 bp, 2% soft-masked lowercase runs.  Returns (names, list of uint8 ASCII arrays)."""
    rng = np.random.default_rng(seed)
    names = ["chr%d" % i for i in range(1, 23)] + ["chrX", "chrY"]
    per = int(total_mbp * 1e6 / len(names))
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs = []
    for _ in names:
        a = lut[rng.integers(0, 4, per, dtype=np.uint8)]
        tel = min(10000, per // 20)
        a[:tel] = ord("N")
        a[-tel:] = ord("N")
        n_target = int(0.005 * per)
        while n_target > 0:
            ln = int(rng.integers(50, 5001))
            s = int(rng.integers(tel, per - tel - ln))
            a[s:s + ln] = ord("N")
            n_target -= ln
        m_target = int(0.02 * per)
        while m_target > 0:
            ln = int(rng.integers(100, 3001))
            s = int(rng.integers(tel, per - tel - ln))
            a[s:s + ln] |= 0x20  # lowercase
            m_target -= ln
        seqs.append(a)
    return names, seqs


def synth_snvs(names, seqs, n, seed):
    """Uniform positions >= 500 bp from contig ends; ref = genome base upper-cased (1% deliberately
    wrong), alt uniform over the other three.  Returns (chrom idx int32, pos int64 1-based, ref, alt codes)."""
    rng = np.random.default_rng(seed)
    code = np.full(256, 4, dtype=np.uint8)
    for i, b in enumerate(b"ACGT"):
        code[b] = i
        code[b | 0x20] = i
    chrom = rng.integers(0, len(names), n).astype(np.int32)
    per = len(seqs[0])
    pos0 = rng.integers(500, per - 500, n)            # 0-based variant position
    ref = np.empty(n, dtype=np.uint8)
    for c in range(len(names)):
        sel = chrom == c
        ref[sel] = code[seqs[c][pos0[sel]]]
    unknown = ref == 4
    ref[unknown] = rng.integers(0, 4, int(unknown.sum()))  # VCF ref is a real base even over N
    wrong = rng.random(n) < 0.01
    ref[wrong] = (ref[wrong] + rng.integers(1, 4, int(wrong.sum()))) % 4
    alt = ((ref + rng.integers(1, 4, n)) % 4).astype(np.uint8)
    return chrom, (pos0 + 1).astype(np.int64), ref, alt


def snv_block(names, seqs, block_size, rank, block):
    """Block ``block`` of rank ``rank``'s variant stream (the reference arm takes prefixes of the same blocks)."""
    return synth_snvs(names, seqs, block_size, [STREAM_SEED, int(rank), int(block)])


class _SeqView(object):
    """str-like view of a uint8 chromosome for the oracle (len + slicing)."""

    def __init__(self, arr):
        self.arr = arr

    def __len__(self):
        return len(self.arr)

    def __getitem__(self, sl):
        return self.arr[sl].tobytes().decode("latin-1")


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler(object):
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            c = [x.strip() for x in s.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py touches oracle/ and baseline/)
# --------------------------------------------------------------------------------------------------
def cpu_port_rate(names, seqs, chrom, pos, ref, alt):
    """The oracle port (numpy encode + torch-CPU DeepSEA batch 64 + scipy scores) on the given variants."""
    import torch
    from oracle import selene_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    layers = O.deepsea_layers(SEQ_LEN, N_FEATURES, seed=0)
    genome = {n: _SeqView(s) for n, s in zip(names, seqs)}
    variants = [(names[chrom[i]], int(pos[i]), "ACGT"[ref[i]], "ACGT"[alt[i]], "+") for i in range(len(pos))]
    t0 = time.perf_counter()
    O.vep_scores(genome, variants, layers, SEQ_LEN, batch_size=64)
    dt = time.perf_counter() - t0
    return 2.0 * len(pos) / dt, dt, torch.get_num_threads()


def make_reference(names, seqs, workdir):
    """ReferenceVEP over the same genome, or None when baseline/_ref is not installed."""
    from baseline import reference_arm as R
    if not R.available():
        return None
    return R.ReferenceVEP(names, seqs, workdir, SEQ_LEN, N_FEATURES, seed=0, threads=os.cpu_count() or 1)


def workload_config(args):
    return {"workload": "DeepSEA variant_effect_prediction (configs[1]): synthetic SNVs on a synthetic "
                        "%d Mbp / 24-contig genome, abs_diffs+logits" % int(args.genome_mbp),
            "variants_per_step_per_gpu": args.variants_per_step, "predictions_per_variant": 2,
            "sequence_length": SEQ_LEN, "n_features": N_FEATURES, "genome_mbp": args.genome_mbp,
            "variant_stream": "block b of rank r = synth_snvs(seed=[2024, r, b]), %d variants per block"
                              % args.variants_per_step,
            "sharding": "independent variants per rank",
            "l2": "per-step working set (activations %.1f GB) >> 126 MB L2; no explicit flush"
                  % (2 * args.variants_per_step * (248 * 320 + 60 * 480 + 53 * 960) * 2 / 1e9)}


def reference_arm(args, rank):
    """--impl reference: the reference's own CPU implementation, all host threads, bounded sample per step."""
    if rank != 0:
        return 0
    names, seqs = synth_genome_arrays(args.genome_mbp)
    per_step = args.ref_variants_per_step
    work = tempfile.mkdtemp(prefix="sb_refarm_")
    ref = make_reference(names, seqs, work)
    kind = "reference" if ref is not None else "port"
    n_total = args.steps + args.warmup
    blocks = [tuple(a[:per_step] for a in snv_block(names, seqs, args.variants_per_step, 0, b)) for b in range(n_total)]

    def step(b):
        c, p, r, a = blocks[b]
        if ref is not None:
            out = ref.score(c, p, r, a)
            assert out["abs_diffs"].shape == (per_step, N_FEATURES) and out["logits"].shape == (per_step, N_FEATURES)
        else:
            cpu_port_rate(names, seqs, c, p, r, a)
    for b in range(args.warmup):
        step(b)
    t0 = time.perf_counter()
    for b in range(args.warmup, n_total):
        step(b)
    dt = time.perf_counter() - t0
    rate = 2.0 * per_step * args.steps / dt
    threads = ref.threads if ref is not None else (os.cpu_count() or 1)
    sample = ("each step = the first %d variants of the corresponding %d-variant block of the same stream, through "
              "%s, file writing excluded" % (per_step, args.variants_per_step,
              "the unmodified selene_sdk AnalyzeSequences.variant_effect_prediction (baseline/_ref, use_cuda=False, "
              "batch 64; its TSV append call replaced by an in-memory collector)" if ref is not None else
              "oracle.vep_scores (baseline/_ref is not installed)"))
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------------
# HBM-bound kernels, timed alone (roofline_hbm)
# --------------------------------------------------------------------------------------------------
def _timed(fn, warmup=3, iters=6):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def hbm_rooflines(G, names, seqs, hbm_gbs):
    """Achieved GB/s = algorithmic bytes per launch / CUDA-event time per launch, inputs larger
    than L2 (126 MB) except where noted."""
    import ctypes as C
    import torch
    from selene_b200 import _lib
    from selene_b200.predict._in_silico_mutagenesis import ism_mutations_device, expand_mutants_device
    lib = _lib.load()
    per = len(seqs[0])
    rng = np.random.default_rng(1)
    out = []

    def emit(kernel, bytes_launch, t, **kw):
        gbs = bytes_launch / t / 1e9
        out.append(dict(kernel=kernel, achieved=gbs, peak=hbm_gbs, unit="GB/s", frac=gbs / hbm_gbs, ms=t * 1e3,
                        algorithmic_bytes_per_launch=bytes_launch, **kw))
    # context: what a plain device fill (write-only stream, torch's vectorised fill kernel) reaches on this GPU
    fill_buf = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    t = _timed(lambda: fill_buf.zero_())
    emit("torch.zero_ (library fill, context only)", float(fill_buf.numel()), t, units="1 GiB", bytes_per_unit=1.0)
    del fill_buf
    n = 65536
    chrom = torch.from_numpy(rng.integers(0, len(names), n).astype(np.int32)).cuda()
    start = torch.from_numpy(rng.integers(0, per - 1000, n).astype(np.int64)).cuda()
    rc = torch.from_numpy(rng.integers(0, 2, n).astype(np.uint8)).cuda()
    for dt, bpw in (("float32", 16375.0), ("bfloat16", 8375.0)):
        t = _timed(lambda: G.encode_windows(chrom, start, 1000, strands=rc, dtype=dt))
        emit("window_kernel[%s]" % dt, n * bpw, t, units="%d windows of 1000 bp" % n, bytes_per_unit=bpw)
    nv = 1 << 18
    vch = torch.from_numpy(rng.integers(0, len(names), nv).astype(np.int32)).cuda()
    vst = torch.from_numpy(rng.integers(0, per - 1000, nv).astype(np.int64)).cuda()
    vref = torch.from_numpy(rng.integers(0, 4, nv).astype(np.uint8)).cuda()
    valt = torch.from_numpy(rng.integers(0, 4, nv).astype(np.uint8)).cuda()
    t = _timed(lambda: G.snv_pairs(vch, vst, vref, valt, 1000, 499, want_onehot=False, want_codes=True))
    emit("snv_pairs_kernel[codes]", nv * 1502.0, t, units="%d variants" % nv, bytes_per_unit=1502.0)
    r0 = np.random.default_rng(0)
    seq = "".join(np.array(list("ACGT"))[r0.integers(0, 4, 1000)].tolist())
    pos, alt, codes = ism_mutations_device(seq)
    pos_r, alt_r = pos.repeat(8), alt.repeat(8)
    t = _timed(lambda: expand_mutants_device(codes, pos_r, alt_r))
    emit("ism_expand_kernel[float32]", pos_r.numel() * 16000.0, t, units="%d mutants" % pos_r.numel(), bytes_per_unit=16000.0)
    F, nrows = N_FEATURES, 2000000
    lr = np.random.default_rng(7)
    c_i = lr.integers(0, len(names), nrows).astype(np.int32)
    s_i = lr.integers(0, per - 3000, nrows).astype(np.int32)
    ln = np.clip(lr.lognormal(np.log(300), 0.8, nrows), 50, 2000).astype(np.int32)
    f_i = lr.integers(0, F, nrows).astype(np.int32)
    h = C.c_void_p()
    e_i = np.ascontiguousarray(s_i + ln)
    _lib.check(lib.sb_intervals_create(_lib.ptr(c_i), _lib.ptr(s_i), _lib.ptr(e_i), _lib.ptr(f_i), nrows, F,
                                       len(names), 0, C.byref(h)), "sb_intervals_create")
    nq = 1 << 19
    qc = torch.from_numpy(lr.integers(0, len(names), nq).astype(np.int32)).cuda()
    qs_np = lr.integers(0, per - 200, nq).astype(np.int32)
    qs, qe = torch.from_numpy(qs_np).cuda(), torch.from_numpy(qs_np + 200).cuda()
    thr = torch.full((F,), 99, dtype=torch.int32, device="cuda")
    hits = float(nrows) * 500.0 / (per * len(names))     # expected overlapping intervals per 200-bp bin
    for dt, code, tdt, sz in (("uint8", _lib.SB_U8, torch.uint8, 1), ("float32", _lib.SB_F32, torch.float32, 4),
                              ("int64", _lib.SB_I64, torch.int64, 8)):
        o = torch.empty((nq, F), dtype=tdt, device="cuda")
        t = _timed(lambda: _lib.check(lib.sb_label_bins(h, _lib.ptr(qc), _lib.ptr(qs), _lib.ptr(qe), nq, _lib.ptr(thr),
                                                        code, _lib.ptr(o), _lib.stream_ptr()), "sb_label_bins"))
        bq = F * sz + 12.0 * hits + 8.0
        emit("label_bins_kernel[%s]" % dt, nq * bq, t, units="%d bins x %d features" % (nq, F), bytes_per_unit=bq)
        del o
    lib.sb_intervals_destroy(h)
    return out


# --------------------------------------------------------------------------------------------------
# strong scaling: a fixed variant list split over the ranks (configs[1])
# --------------------------------------------------------------------------------------------------
def strong_leg(az, names, seqs, args, rank, world):
    import torch
    import torch.distributed as dist
    from selene_b200.sharding import shard_range, gather_tiles
    M, V = args.strong_variants, args.variants_per_step
    lo, hi = shard_range(M, rank, world)
    n_local = hi - lo
    # the global list is rank 0's block stream; this rank materialises the blocks that cover [lo, hi)
    parts = []
    for b in range(lo // V, (hi + V - 1) // V):
        blk = snv_block(names, seqs, V, 0, 1000 + b)
        a, z = max(lo, b * V) - b * V, min(hi, (b + 1) * V) - b * V
        parts.append(tuple(x[a:z] for x in blk))
    chrom, pos, ref, alt = (np.concatenate([p[i] for p in parts]) for i in range(4))
    save = ("abs_diffs", "logits")
    tiles = {k: torch.empty((n_local, N_FEATURES), dtype=torch.float32, device="cuda") for k in save}
    host = {k: torch.empty((n_local, N_FEATURES), dtype=torch.float32).pin_memory() for k in save}
    copy_stream = torch.cuda.Stream()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def one_pass(read_back):
        for a in range(0, n_local, V):
            z = min(n_local, a + V)
            prep = az.prepare_snvs(chrom[a:z], pos[a:z], ref[a:z], alt[a:z], pinned=True)
            az.score_prepared(prep, save, outs={k: tiles[k][a:z] for k in save})
            if read_back:
                done = torch.cuda.Event()
                done.record()
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(done)
                    for k in save:
                        host[k][a:z].copy_(tiles[k][a:z], non_blocking=True)

    one_pass(False)                                            # warm-up (allocator, workspaces)
    sync_all()
    t0 = time.perf_counter(); one_pass(False); sync_all()
    compute_s = max_over_ranks(time.perf_counter() - t0)
    t0 = time.perf_counter(); one_pass(True); sync_all()
    shards_s = max_over_ranks(time.perf_counter() - t0)
    # (b) one NCCL gather of the score tiles to rank 0, then rank 0 reads the whole matrix back
    gather_bytes = 0
    full = {k: gather_tiles(tiles[k], M, dst=0) for k in save}   # untimed: NCCL channel set-up, receive buffers
    sync_all()
    t0 = time.perf_counter()
    full = {k: gather_tiles(tiles[k], M, dst=0, out=full[k] if rank == 0 else None) for k in save}
    sync_all()
    gather_s = max_over_ranks(time.perf_counter() - t0)
    if world > 1:
        gather_bytes = int((M - n_local) * N_FEATURES * 4 * len(save)) if rank == 0 else 0
    d2h0_s = 0.0
    if rank == 0:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in save:                                          # through the pinned shard buffers, n_local rows at a time
            for a in range(0, M, max(n_local, 1)):
                z = min(M, a + n_local)
                host[k][:z - a].copy_(full[k][a:z], non_blocking=True)
        torch.cuda.synchronize()
        d2h0_s = time.perf_counter() - t0
    d2h0_s = max_over_ranks(d2h0_s)
    gather_bytes = int(max_over_ranks(float(gather_bytes)))
    gather_path = compute_s + gather_s + d2h0_s
    best = min(gather_path, shards_s)
    return {"variants": M, "n_gpus": world, "scaling": "strong", "compute_s": compute_s,
            "shards_path_s": shards_s, "gather_s": gather_s, "gather_bytes": gather_bytes, "rank0_d2h_s": d2h0_s,
            "gather_path_s": gather_path, "variants_per_s": M / best, "predictions_per_s": 2.0 * M / best,
            "default": "shards" if shards_s <= gather_path else "gather",
            "what": "shards_path = H2D of the shard's variant arrays + scoring + device->host read of the shard's score "
                    "tiles on a copy stream (what the per-rank TSV writers consume); gather_path = scoring + one NCCL "
                    "gather of the (n_local, 919) abs_diffs and logits tiles to rank 0 over NVLink + rank 0's "
                    "device->host read of the whole (variants, 919) x 2 matrix over its one PCIe link"}


# --------------------------------------------------------------------------------------------------
# configs[2]: DeepSEA training step, data parallel (secondary record; weak scaling, batch 64 per GPU)
# --------------------------------------------------------------------------------------------------
def train_leg(model_cls, rank, world, batch=64, steps=20, warmup=5, genome=None, names=None, seqs=None):
    """forward (train mode, dropout) -> BCELoss -> backward on tcgen05 (bf16 operands, fp32 master weights) -> NCCL
    all-reduce of the flat gradient vector as bf16 (classifier tail overlapped with the conv backward) -> SGD, on a fixed
    synthetic batch per rank (selene_sdk/train_model.py:483-496, models/deepsea.py:60-74); then the same step fed by
    RandomPositionsSampler + GenomicFeatures labelling (919 features, 200-bp bins) with the next batch drawn while the
    GPU steps (selene_sdk/samplers/random_positions_sampler.py:264-369)."""
    import torch
    import torch.distributed as dist
    from selene_b200.train_model import B200Trainer
    torch.manual_seed(0)
    model = model_cls()
    tr = B200Trainer(model, SEQ_LEN, lr=0.01, momentum=0.9, weight_decay=1e-6,
                     dropout=B200Trainer.dropout_of_module(model), precision="bf16")
    rng = np.random.default_rng(1 + rank)
    x = np.eye(4, dtype=np.float32)[rng.integers(0, 4, (batch, SEQ_LEN))]
    y = (rng.random((batch, N_FEATURES)) < 0.1).astype(np.float32)
    xs, ys = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    for _ in range(warmup):
        tr.step(xs, ys)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        tr._enable_overlap()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        tr.forward_backward(xs, ys)
        tr.allreduce_grads(overlap=True, sgd_tail=True)
        tr.sgd_step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    out = {"workload": "configs[2]: DeepSEA training step (dropout, BCELoss, SGD momentum 0.9), batch %d per GPU, "
                       "synthetic fixed batch (no sampler)" % batch,
           "n_gpus": world, "scaling": "weak", "precision": "bf16 operands / fp32 master weights",
           "ms_per_step": ms, "samples_per_s": world * batch / ms * 1e3,
           "allreduce_bytes_per_step": int(tr.grads.numel()) * (2 if tr.grad_comm_dtype == "bf16" else 4) if world > 1 else 0,
           "grad_comm_dtype": tr.grad_comm_dtype,
           "flops_per_sample": 3.0 * 1.0986e9, "loss": float(tr.loss.item())}
    out["tflops"] = out["samples_per_s"] * out["flops_per_sample"] / 1e12
    if genome is not None:
        try:
            out["with_sampler"] = _train_with_sampler(tr, genome, names, seqs, rank, world, batch, steps)
        except Exception as e:      # secondary record: never lose the bench line to it
            out["with_sampler"] = {"error": "%s: %s" % (type(e).__name__, e)}
    del tr
    torch.cuda.empty_cache()
    return out


def _train_with_sampler(tr, genome, names, seqs, rank, world, batch, steps):
    """configs[2] end to end: every step's batch is drawn by RandomPositionsSampler (encode + label kernels) while the
    previous step runs on the GPU; the loss of every step is read on the host."""
    import gzip
    import tempfile
    import torch
    import torch.distributed as dist
    from selene_b200.samplers import RandomPositionsSampler
    per = len(seqs[0])
    rng = np.random.default_rng(11)
    n_iv = 300000
    c_i = rng.integers(0, len(names), n_iv)
    s_i = rng.integers(0, per - 3000, n_iv)
    ln = np.clip(rng.lognormal(np.log(300), 0.8, n_iv), 50, 2000).astype(np.int64)
    f_i = rng.integers(0, N_FEATURES, n_iv)
    order = np.lexsort((s_i, c_i))
    feats = ["f%d" % i for i in range(N_FEATURES)]
    td = tempfile.TemporaryDirectory()       # lives until the function returns
    bed = os.path.join(td.name, "features_rank%d.bed.gz" % rank)
    with gzip.open(bed, "wt", compresslevel=1) as fh:
        fh.write("".join("%s\t%d\t%d\tf%d\n" % (names[c_i[k]], s_i[k], s_i[k] + ln[k], f_i[k]) for k in order))
    smp = RandomPositionsSampler(genome, bed, feats, seed=436 + rank, validation_holdout=[names[-1]],
                                 test_holdout=[names[-2]], sequence_length=SEQ_LEN, center_bin_to_predict=200,
                                 feature_thresholds=0.5)
    draw = lambda: smp.sample_device(batch, seq_dtype="float32", label_dtype="float32")
    nxt = draw()
    for _ in range(3):
        tr.step_async(*nxt)
        nxt = draw()
        tr.loss_value()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.step_async(*nxt)
        nxt = draw()
        last = tr.loss_value()
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item()) / steps
    td.cleanup()
    return {"what": "RandomPositionsSampler.sample_device(%d) (919 features, 200-bp bins, 300 k intervals) -> step; the next "
                    "batch is drawn on the host while the GPU steps; loss read every step; wall clock" % batch,
            "ms_per_step": dt * 1e3, "samples_per_s": world * batch / dt, "loss": last}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variants-per-step", type=int, default=16384)
    ap.add_argument("--ref-variants-per-step", type=int, default=1024,
                    help="reference arm: variants of every block actually run on the CPU (bounded sample)")
    ap.add_argument("--genome-mbp", type=float, default=240.0)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-sample", type=int, default=8192,
                    help="variants timed on the CPU baseline leg (reference: ~10 s on 16 threads; the oracle port runs on the first 1024)")
    ap.add_argument("--strong-variants", type=int, default=1 << 20, help="0 skips the strong-scaling leg")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-hbm", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return reference_arm(args, rank)
    config = workload_config(args)

    # ---------------------------------------------------------------- our arm (GPU)
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from selene_b200 import _lib
    from selene_b200.sequences import Genome
    from selene_b200.predict.model_predict import AnalyzeSequences
    lib = _lib.load()

    names, seqs = synth_genome_arrays(args.genome_mbp)
    genome = Genome.from_sequences(names, seqs, device=local_rank)
    import torch.nn as nn

    class DeepSEA(nn.Module):  # the reference architecture, models/deepsea.py:21-50, random-init
        def __init__(self):
            super().__init__()
            self.conv_net = nn.Sequential(
                nn.Conv1d(4, 320, 8), nn.ReLU(inplace=True), nn.MaxPool1d(4, 4), nn.Dropout(p=0.2),
                nn.Conv1d(320, 480, 8), nn.ReLU(inplace=True), nn.MaxPool1d(4, 4), nn.Dropout(p=0.2),
                nn.Conv1d(480, 960, 8), nn.ReLU(inplace=True), nn.Dropout(p=0.5))
            self.classifier = nn.Sequential(nn.Linear(960 * 53, N_FEATURES), nn.ReLU(inplace=True),
                                            nn.Linear(N_FEATURES, N_FEATURES), nn.Sigmoid())
    torch.manual_seed(0)
    model = DeepSEA()
    az = AnalyzeSequences(model, sequence_length=SEQ_LEN, features=["f%d" % i for i in range(N_FEATURES)],
                          batch_size=64, reference_sequence=genome, precision=args.precision, device=local_rank)

    V = args.variants_per_step
    n_steps_total = args.steps + args.warmup
    blocks = [snv_block(names, seqs, V, rank, b) for b in range(n_steps_total)]
    save = ("abs_diffs", "logits")
    outs = {k: torch.empty((V, N_FEATURES), dtype=torch.float32, device="cuda") for k in save}

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- leg 1: device-resident inputs (value)
    preps = [az.prepare_snvs(*blocks[i]) for i in range(n_steps_total)]
    for i in range(args.warmup):
        az.score_prepared(preps[i], save, outs=outs)
    sync_all()
    lib.sb_net_set_timing(az.net._h, 1)
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = lib.sb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.warmup, n_steps_total):
        az.score_prepared(preps[i], save, outs=outs)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = int(lib.sb_launch_count() - launches0)
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = 2.0 * V * args.steps * world / (ms_max * 1e-3)

    # per-layer device times -> roofline of the dominant kernel
    import ctypes as C

    def layer_times():
        ms_, n_ = [], []
        for li in range(lib.sb_net_n_layers(az.net._h)):
            tm, ln = C.c_double(), C.c_int()
            _lib.check(lib.sb_net_layer_time(az.net._h, li, C.byref(tm), C.byref(ln)), "sb_net_layer_time")
            ms_.append(tm.value); n_.append(ln.value)
        return ms_, n_
    layer_ms_ovl, layer_n_ovl = layer_times()
    lib.sb_net_set_timing(az.net._h, 0)
    # In the timed region above the first conv of chunk k+1 runs on a side stream NEXT TO the tcgen05 layers of chunk k
    # (same SMs), so a layer's event time there is the time of two kernels sharing the machine.  The roofline of the
    # dominant kernel is taken from a second timed pass over the same blocks with that overlap switched off
    # (SB_CONV1_AHEAD=0: every kernel alone on its stream); both sets of layer times are reported.
    alone_steps = min(args.steps, 4)
    os.environ["SB_CONV1_AHEAD"] = "0"
    az.score_prepared(preps[args.warmup], save, outs=outs)
    sync_all()
    lib.sb_net_set_timing(az.net._h, 1)
    ea0, ea1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea0.record()
    for i in range(args.warmup, args.warmup + alone_steps):
        az.score_prepared(preps[i], save, outs=outs)
    ea1.record()
    sync_all()
    ms_alone = ea0.elapsed_time(ea1) / alone_steps
    layer_ms, layer_n = layer_times()
    lib.sb_net_set_timing(az.net._h, 0)
    del os.environ["SB_CONV1_AHEAD"]
    dom = int(np.argmax(layer_ms))
    rows_per_launch = 2.0 * V * alone_steps / max(layer_n[dom], 1)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained"
    hbm_gbs = float(peaks.get("hbm_gbs", 6650.0))
    flops_launch = lib.sb_net_layer_flops_per_seq(az.net._h, dom) * rows_per_launch
    achieved_tf = flops_launch / (layer_ms[dom] / max(layer_n[dom], 1) * 1e-3) / 1e12
    layer_names = ["conv1(lut)", "conv2", "conv3", "fc1", "fc2"]
    # traffic: dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel from the round's
    # committed `ncu --set full` capture (profiles/traffic.json, written by scripts/summarize_ncu.py); null without it
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tj.get(layer_names[dom] if dom < 5 else str(dom))
        if ent and abs(float(ent["rows_per_launch"]) / rows_per_launch - 1.0) < 0.05:
            # the capture is one full chunk; the step's last chunk is ragged, so the average launch has slightly fewer
            # rows: DRAM bytes scale with the rows (activations in / out dominate, weights are 0.3 %)
            traffic = float(ent["dram_bytes"]) * rows_per_launch / float(ent["rows_per_launch"])
            traffic_src = ent["source"] + " (scaled %d -> %.0f rows per launch)" % (int(ent["rows_per_launch"]), rows_per_launch)
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": "tc_layer_kernel[%s]" % layer_names[dom] if dom < 5 else str(dom),
                "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "flops_per_launch": flops_launch, "rows_per_launch": rows_per_launch,
                "timing": "dominant kernel alone on its stream: second timed pass of %d steps over the same blocks with the "
                          "first conv's side-stream overlap off (%.2f ms per step); frac_in_timed_region is the same "
                          "kernel while it shares the SMs with the next chunk's first conv" % (alone_steps, ms_alone),
                "frac_in_timed_region": (lib.sb_net_layer_flops_per_seq(az.net._h, dom) * 2.0 * V * args.steps /
                                         (layer_ms_ovl[dom] * 1e-3) / 1e12) / peak_tf,
                "layer_ms_per_step": {layer_names[i] if i < 5 else str(i): layer_ms[i] / alone_steps
                                      for i in range(len(layer_ms))},
                "layer_ms_per_step_in_timed_region": {layer_names[i] if i < 5 else str(i): layer_ms_ovl[i] / args.steps
                                                      for i in range(len(layer_ms_ovl))},
                "whole_net_tflops": az.net.flops_per_seq * 2.0 * V * args.steps / (ms_max * 1e-3) / 1e12}
    roofline["whole_net_frac"] = roofline["whole_net_tflops"] / peak_tf

    # ---- leg 2: end to end from host buffers through the public API (H2D + D2H inside the timed region)
    # Two sets of device / pinned host result buffers: the device->host read of step i runs on a copy stream while
    # step i+1 computes (every step still copies its own inputs in and its own 120 MB of scores out).
    host_out = [{k: torch.empty((V, N_FEATURES), dtype=torch.float32).pin_memory() for k in save} for _ in range(2)]
    dev_out = [outs, {k: torch.empty_like(v) for k, v in outs.items()}]
    d2h_bytes = sum(v.numel() * 4 for v in host_out[0].values())
    h2d_bytes = 0
    copy_stream = torch.cuda.Stream()
    copy_done = [None, None]

    def e2e_step(i):
        nonlocal h2d_bytes
        b = i & 1
        if copy_done[b] is not None:
            copy_done[b].synchronize()          # buffer set b: its previous read-back has landed on the host
        prep = az.prepare_snvs(*blocks[i], pinned=True)
        h2d_bytes = prep["h2d_bytes"]
        r = az.score_prepared(prep, save, outs=dev_out[b])
        computed = torch.cuda.Event()
        computed.record()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(computed)
            for k in save:
                host_out[b][k].copy_(r[k], non_blocking=True)
            copy_done[b] = torch.cuda.Event()
            copy_done[b].record(copy_stream)
    for i in range(min(args.warmup, 2)):
        e2e_step(i)
    sync_all()
    t0 = time.perf_counter()
    for i in range(args.warmup, n_steps_total):
        e2e_step(i)
    sync_all()
    dt = time.perf_counter() - t0
    clk = clocks.stop()
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = 2.0 * V * args.steps * world / float(t.item())
    del host_out, dev_out

    # ---- strong scaling of the fixed variant list (configs[1])
    strong = None
    if args.strong_variants > 0:
        strong = strong_leg(az, names, seqs, args, rank, world)

    # ---- configs[2]: the training step, data parallel
    train = None
    if not args.no_train and args.precision == "bf16":
        train = train_leg(DeepSEA, rank, world, genome=genome, names=names, seqs=seqs)

    # ---- HBM-bound kernels alone (N=1)
    hbm = None
    if rank == 0 and world == 1 and not args.no_hbm:
        hbm = hbm_rooflines(genome, names, seqs, hbm_gbs)

    # ---- CPU baseline (rank 0, N=1 only): the reference itself on a prefix of the first timed block
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n_cpu = min(args.cpu_sample, V)
        c, p, r, a = (x[:n_cpu] for x in blocks[args.warmup])
        n_port = min(1024, n_cpu)
        port_rate, port_dt, threads = cpu_port_rate(names, seqs, c[:n_port], p[:n_port], r[:n_port], a[:n_port])
        work = tempfile.mkdtemp(prefix="sb_cpuleg_")
        ref = make_reference(names, seqs, work)
        if ref is not None:
            ref.score(c[:64], p[:64], r[:64], a[:64])           # warm-up batch
            t0 = time.perf_counter()
            ref.score(c, p, r, a)
            rdt = time.perf_counter() - t0
            cpu = {"value": 2.0 * n_cpu / rdt, "unit": UNIT, "cores": ref.threads, "kind": "reference",
                   "sample": "first %d SNVs of the first timed block (%d forwards, %.1f s) through the unmodified "
                             "selene_sdk AnalyzeSequences.variant_effect_prediction (baseline/_ref, use_cuda=False, batch "
                             "64, TSV append replaced by an in-memory collector)" % (n_cpu, 2 * n_cpu, rdt),
                   "port_value": port_rate, "port_seconds": port_dt}
        else:
            cpu = {"value": port_rate, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "first %d SNVs of the first timed block (%d forwards, %.1f s) through oracle.vep_scores "
                             "(baseline/_ref not installed)" % (n_port, 2 * n_port, port_dt)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                "config": config, "clocks": clk, "gpu_launches": launches,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": d2h_bytes},
                "roofline": roofline, "roofline_hbm": hbm, "strong": strong, "train": train, "cpu_baseline": cpu,
                "variants_per_sec": value / 2.0}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
