"""bench.py — measures BASELINE.json's metric, "mutant+variant predictions/sec (DeepSEA 1kb, 919 tgt)",
on the workload of configs[1]: DeepSEA variant_effect_prediction over synthetic SNVs on a synthetic
genome, abs_diffs + logits scoring.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path (SNV windows -> ref/alt rows -> DeepSEA forward -> scores) over
one chunk of --variants-per-step variants (2 predictions each) per GPU.  Under torchrun every rank
runs its own shard (independent variants, no data-path collective; weak scaling).  One JSON line is
printed by rank 0.  The ``--impl reference`` arm times the CPU restatement of the reference's path
(oracle/, torch CPU with all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mutant+variant predictions/sec (DeepSEA 1kb, 919 tgt)"
UNIT = "predictions/s"
SEQ_LEN, N_FEATURES = 1000, 919


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
                 "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
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
# CPU legs (oracle; the only place bench.py touches oracle/)
# --------------------------------------------------------------------------------------------------
def cpu_vep_sample(names, seqs, chrom, pos, ref, alt, n_variants):
    import torch
    from oracle import selene_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    layers = O.deepsea_layers(SEQ_LEN, N_FEATURES, seed=0)
    genome = {n: _SeqView(s) for n, s in zip(names, seqs)}
    variants = [(names[chrom[i]], int(pos[i]), "ACGT"[ref[i]], "ACGT"[alt[i]], "+") for i in range(n_variants)]
    t0 = time.perf_counter()
    O.vep_scores(genome, variants, layers, SEQ_LEN, batch_size=64)
    dt = time.perf_counter() - t0
    return 2.0 * n_variants / dt, dt, torch.get_num_threads()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variants-per-step", type=int, default=16384)
    ap.add_argument("--genome-mbp", type=float, default=240.0)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-sample", type=int, default=1024, help="variants timed on the CPU baseline leg")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    config = {"workload": "DeepSEA variant_effect_prediction (configs[1]): synthetic SNVs on a synthetic "
                          "%d Mbp / 24-contig genome, abs_diffs+logits" % int(args.genome_mbp if args.impl == "ours" else min(args.genome_mbp, 24.0)),
              "variants_per_step_per_gpu": args.variants_per_step, "predictions_per_variant": 2,
              "sequence_length": SEQ_LEN, "n_features": N_FEATURES, "sharding": "independent variants per rank"}

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        names, seqs = synth_genome_arrays(min(args.genome_mbp, 24.0))
        per_step = 128
        total = per_step * (args.steps + args.warmup)
        chrom, pos, ref, alt = synth_snvs(names, seqs, total, 2024)
        rate_w = None
        if args.warmup:
            n_w = per_step * args.warmup
            rate_w = cpu_vep_sample(names, seqs, chrom, pos, ref, alt, min(n_w, per_step))
        t0 = time.perf_counter()
        sl = slice(per_step * args.warmup, total)
        rate, dt, threads = cpu_vep_sample(names, seqs, chrom[sl], pos[sl], ref[sl], alt[sl], per_step * args.steps)
        line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": dict(config, variants_per_step_per_gpu=per_step),
                "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                 "sample": "%d SNVs per step x %d steps through oracle.vep_scores (torch CPU, "
                                           "batch 64), file writing excluded" % (per_step, args.steps)},
                "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

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
    chrom, pos, ref, alt = synth_snvs(names, seqs, V * n_steps_total, 2024 + rank)
    save = ("abs_diffs", "logits")
    outs = {k: torch.empty((V, N_FEATURES), dtype=torch.float32, device="cuda") for k in save}

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- leg 1: device-resident inputs (value)
    preps = [az.prepare_snvs(chrom[i * V:(i + 1) * V], pos[i * V:(i + 1) * V], ref[i * V:(i + 1) * V],
                             alt[i * V:(i + 1) * V]) for i in range(n_steps_total)]
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
    clk = clocks.stop()
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = 2.0 * V * args.steps * world / (ms_max * 1e-3)

    # per-layer device times -> roofline of the dominant kernel
    import ctypes as C
    layer_ms, layer_n = [], []
    for li in range(lib.sb_net_n_layers(az.net._h)):
        tm, ln = C.c_double(), C.c_int()
        _lib.check(lib.sb_net_layer_time(az.net._h, li, C.byref(tm), C.byref(ln)), "sb_net_layer_time")
        layer_ms.append(tm.value); layer_n.append(ln.value)
    lib.sb_net_set_timing(az.net._h, 0)
    dom = int(np.argmax(layer_ms))
    rows_per_launch = 2.0 * V * args.steps / max(layer_n[dom], 1)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained"
    flops_launch = lib.sb_net_layer_flops_per_seq(az.net._h, dom) * rows_per_launch
    achieved_tf = flops_launch / (layer_ms[dom] / max(layer_n[dom], 1) * 1e-3) / 1e12
    layer_names = ["conv1(lut)", "conv2", "conv3", "fc1", "fc2"]
    roofline = {"bound": "tensor", "kernel": "tc_layer_kernel[%s]" % layer_names[dom] if dom < 5 else str(dom),
                "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                # dram__bytes_read+write of one conv2 launch of 8192 rows (ncu --set full, profiles/r01a_summary.md);
                # algorithmic bytes of that launch: 8192*(248*320 + 60*480)*2 + weights 2.5 MB = 1.77 GB
                "traffic": 1.757e9 if dom == 1 else None, "peak_source": peak_src,
                "layer_ms_per_step": {layer_names[i] if i < 5 else str(i): layer_ms[i] / args.steps
                                      for i in range(len(layer_ms))},
                "whole_net_tflops": az.net.flops_per_seq * 2.0 * V * args.steps / (ms_max * 1e-3) / 1e12}

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
        sl = slice(i * V, (i + 1) * V)
        prep = az.prepare_snvs(chrom[sl], pos[sl], ref[sl], alt[sl], pinned=True)
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
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = 2.0 * V * args.steps * world / float(t.item())

    # ---- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, cdt, threads = cpu_vep_sample(names, seqs, chrom, pos, ref, alt, min(args.cpu_sample, len(pos)))
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "first %d SNVs of the step stream (%d forwards, %.1f s) through oracle.vep_scores: "
                         "numpy encode + torch-CPU DeepSEA batch 64 + numpy/scipy scores, file writing excluded"
                         % (min(args.cpu_sample, len(pos)), 2 * min(args.cpu_sample, len(pos)), cdt)}

    if rank == 0:
        config.update({"l2": "per-step working set (activations %.1f GB) >> 126 MB L2; no explicit flush"
                             % (2 * V * (248 * 320 + 60 * 480 + 53 * 960) * 2 / 1e9),
                       "precision": args.precision, "genome_mbp": args.genome_mbp})
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                "config": config, "clocks": clk, "gpu_launches": launches,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": d2h_bytes},
                "roofline": roofline, "cpu_baseline": cpu, "variants_per_sec": value / 2.0}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
