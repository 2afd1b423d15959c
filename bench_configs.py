"""Secondary measurements (not the driver's contract — that is bench.py): the HBM-bound kernels
against the measured copy bandwidth, and BASELINE.json's other configs (ISM latency, DanQ ISM,
Seqweaver / HeartENN sweeps, sampler rate).  Prints one JSON object per line; run on a B200:

    python bench_configs.py [--quick]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def timed(fn, warmup=3, iters=10):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="", help="comma list of sections: encode,labels,ism,vepfile,danq,sweep,sampler,train")
    args = ap.parse_args()
    only = set(x for x in args.only.split(",") if x)
    want = lambda name: not only or name in only
    import torch
    import bench as B
    import sb_helpers as H
    import sb_models as M
    from selene_b200 import _lib
    from selene_b200.sequences import Genome
    from selene_b200.targets import GenomicFeatures
    from selene_b200.nets import B200Net, layers_from_module
    from selene_b200.predict.model_predict import AnalyzeSequences
    from selene_b200.predict._in_silico_mutagenesis import ism_mutations_device, expand_mutants_device
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
    hbm = float(peaks["hbm_gbs"])
    out = []

    def emit(**kw):
        print(json.dumps(kw))
        sys.stdout.flush()
        out.append(kw)

    names, seqs = B.synth_genome_arrays(48.0 if args.quick else 240.0)
    G = Genome.from_sequences(names, seqs)
    Genome.update_bases_order(['A', 'C', 'G', 'T'])
    per = len(seqs[0])
    rng = np.random.default_rng(1)

    # ---------------------------------------------------------------- kernel (a): encode windows
    n = 65536 if want('encode') else 0
    chrom = torch.from_numpy(rng.integers(0, len(names), n).astype(np.int32)).cuda()
    start = torch.from_numpy(rng.integers(0, per - 1000, n).astype(np.int64)).cuda()
    rc = torch.from_numpy(rng.integers(0, 2, n).astype(np.uint8)).cuda()
    for dt, bpw in ((("float32", 16375.0), ("bfloat16", 8375.0)) if want('encode') else ()):
        t = timed(lambda: G.encode_windows(chrom, start, 1000, strands=rc, dtype=dt))
        gbs = n * bpw / t / 1e9
        emit(kernel="sb_encode_windows", dtype=dt, windows=n, L=1000, ms=t * 1e3, algorithmic_bytes_per_window=bpw,
             achieved_gbs=gbs, peak_gbs=hbm, frac=gbs / hbm, windows_per_s=n / t)
    # ---------------------------------------------------------------- ISM expansion (one-hot mutants)
    r = np.random.default_rng(0)
    seq = "".join(np.array(list("ACGT"))[r.integers(0, 4, 1000)].tolist())
    pos, alt, codes = ism_mutations_device(seq)
    reps = 8
    pos_r, alt_r = pos.repeat(reps), alt.repeat(reps)
    if want('encode'):
        t = timed(lambda: expand_mutants_device(codes, pos_r, alt_r))
        gbs = pos_r.numel() * 16000.0 / t / 1e9
        emit(kernel="sb_ism_expand", mutants=int(pos_r.numel()), ms=t * 1e3, achieved_gbs=gbs, peak_gbs=hbm, frac=gbs / hbm)

    # ---------------------------------------------------------------- kernel (b): labels
    F = 919
    nrows = 400000 if args.quick else 2000000
    lr = np.random.default_rng(7)
    c_i = lr.integers(0, len(names), nrows).astype(np.int32)
    s_i = lr.integers(0, per - 3000, nrows).astype(np.int32)
    ln = np.clip(lr.lognormal(np.log(300), 0.8, nrows), 50, 2000).astype(np.int32)
    f_i = lr.integers(0, F, nrows).astype(np.int32)
    import ctypes as C
    lib = _lib.load()
    h = C.c_void_p()
    e_i = np.ascontiguousarray(s_i + ln)
    _lib.check(lib.sb_intervals_create(_lib.ptr(c_i), _lib.ptr(s_i), _lib.ptr(e_i),
                                       _lib.ptr(f_i), nrows, F, len(names), 0, C.byref(h)), "sb_intervals_create")
    nq = 65536 if args.quick else 1 << 20   # int64 labels of 2^20 bins = 7.7 GB written per launch
    qc = torch.from_numpy(lr.integers(0, len(names), nq).astype(np.int32)).cuda()
    qs_np = lr.integers(0, per - 200, nq).astype(np.int32)
    qs = torch.from_numpy(qs_np).cuda()
    qe = torch.from_numpy(qs_np + 200).cuda()
    thr = torch.full((F,), 99, dtype=torch.int32, device="cuda")
    for dt, code, tdt, sz in ((("uint8", _lib.SB_U8, torch.uint8, 1), ("float32", _lib.SB_F32, torch.float32, 4),
                               ("int64", _lib.SB_I64, torch.int64, 8)) if want('labels') else ()):
        o = torch.empty((nq, F), dtype=tdt, device="cuda")
        t = timed(lambda: _lib.check(lib.sb_label_bins(h, _lib.ptr(qc), _lib.ptr(qs), _lib.ptr(qe), nq, _lib.ptr(thr),
                                                       code, _lib.ptr(o), _lib.stream_ptr()), "sb_label_bins"))
        hits = float(nrows) * 500.0 / (per * len(names))     # expected overlapping intervals per query
        bytes_q = F * sz + 12.0 * hits
        gbs = nq * bytes_q / t / 1e9
        emit(kernel="sb_label_bins", out_dtype=dt, queries=nq, features=F, intervals=nrows, ms=t * 1e3,
             algorithmic_bytes_per_query=bytes_q, achieved_gbs=gbs, peak_gbs=hbm, frac=gbs / hbm, queries_per_s=nq / t,
             positives=float(o.float().mean().item()))
    lib.sb_intervals_destroy(h)

    # ---------------------------------------------------------------- config 1: DeepSEA ISM of one sequence
    torch.manual_seed(0)
    feats = ["f%d" % i for i in range(919)]
    az = AnalyzeSequences(M.DeepSEA(1000, 919), sequence_length=1000, features=feats, reference_sequence=Genome)
    for prec in (("bf16", "fp32") if want('ism') else ()):
        az.precision = prec
        for _ in range(2):
            az.ism_scores(seq, ("abs_diffs", "logits"))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        k = 5 if prec == "bf16" else 2
        for _ in range(k):
            _, res, _ = az.ism_scores(seq, ("abs_diffs", "logits"))
        dt_ = (time.perf_counter() - t0) / k
        emit(config="1: DeepSEA ISM, one 1000-bp sequence, 3000 mutants, abs_diffs+logits to host", precision=prec,
             seconds=dt_, mutants_per_s=3000 / dt_)

    # ISM with the reference's output files: the "{:.2e}" text is produced on the device
    if want('ism'):
        import io, shutil, tempfile
        az.precision = "bf16"
        tmpd = tempfile.mkdtemp(prefix="sb_ism_")
        az.in_silico_mutagenesis(seq, ["abs_diffs", "logits"], output_path_prefix=os.path.join(tmpd, "w"))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(3):
            az.in_silico_mutagenesis(seq, ["abs_diffs", "logits"], output_path_prefix=os.path.join(tmpd, "r%d" % i))
        dt_files = (time.perf_counter() - t0) / 3
        size = os.path.getsize(os.path.join(tmpd, "r0_abs_diffs.tsv"))
        shutil.rmtree(tmpd, ignore_errors=True)
        m_ = np.random.default_rng(0).random((3000, 919)).astype(np.float32) * 1e-3
        t0 = time.perf_counter()
        b_ = io.StringIO()
        np.savetxt(b_, m_.astype(np.float64), fmt="%.2e", delimiter="\t")
        host_fmt = time.perf_counter() - t0
        emit(config="1: DeepSEA ISM incl. writing abs_diffs + logits TSV files (2 x 3000 x 919 values)", precision="bf16",
             seconds=dt_files, mutants_per_s=3000 / dt_files, tsv_bytes_per_file=size,
             host_savetxt_seconds_per_matrix=host_fmt)

    # ---------------------------------------------------------------- config 2 through the public API with files
    if want('vepfile'):
        import shutil, tempfile
        nv = 20000 if args.quick else 100000
        tmpd = tempfile.mkdtemp(prefix="sb_vep_")
        vcf = os.path.join(tmpd, "v.vcf")
        vr = np.random.default_rng(2024)
        ci_, pos_, alt_ = vr.integers(0, len(names), nv), vr.integers(600, per - 600, nv), vr.integers(0, 4, nv)
        with open(vcf, "w") as fh:
            fh.write("#CHROM\tPOS\tID\tREF\tALT\n")
            for i in range(nv):
                ref_ = chr(seqs[ci_[i]][pos_[i] - 1]).upper()
                fh.write("%s\t%d\trs%d\t%s\t%s\n" % (names[ci_[i]], pos_[i], i, ref_ if ref_ in "ACGT" else "A", "ACGT"[alt_[i]]))
        azv = AnalyzeSequences(M.DeepSEA(1000, 919), sequence_length=1000, features=feats, reference_sequence=G)
        import warnings as _w
        with _w.catch_warnings():
            _w.simplefilter("ignore")
            azv.variant_effect_prediction(vcf, ["abs_diffs", "logits"], output_dir=os.path.join(tmpd, "w"))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            azv.variant_effect_prediction(vcf, ["abs_diffs", "logits"], output_dir=os.path.join(tmpd, "o"))
            torch.cuda.synchronize()
        dtv = time.perf_counter() - t0
        mb = sum(os.path.getsize(os.path.join(tmpd, "o", f)) for f in os.listdir(os.path.join(tmpd, "o"))) / 1e6
        shutil.rmtree(tmpd, ignore_errors=True)
        emit(config="2: AnalyzeSequences.variant_effect_prediction from a VCF file to abs_diffs + logits TSV files",
             precision="bf16", variants=nv, seconds=dtv, variants_per_s=nv / dtv, predictions_per_s=2 * nv / dtv,
             tsv_mb_written=mb)

    # ---------------------------------------------------------------- config 4: DanQ ISM
    torch.manual_seed(0)
    azd = AnalyzeSequences(M.DanQ(1000, 919), sequence_length=1000, features=feats, reference_sequence=Genome)
    n_seq = 4 if args.quick else 8
    if not want('danq'):
        n_seq = 0
    else:
        for _ in range(3):
            azd.ism_scores(seq, ("abs_diffs", "logits"), to_host=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n_seq):
        azd.ism_scores(seq, ("abs_diffs", "logits"), to_host=False)
    torch.cuda.synchronize()
    dt_ = (time.perf_counter() - t0) / max(n_seq, 1)
    if want('danq'):
      emit(config="4: DanQ ISM (batch-axis recurrence, groups of 64), per 1000-bp sequence of 3000 mutants", precision="bf16",
         seconds_per_sequence=dt_, mutants_per_s=3000 / dt_, flops_per_seq=azd.net.flops_per_seq,
         tflops=3000 / dt_ * azd.net.flops_per_seq / 1e12)

    # the same workload with two sequences in flight on two CUDA streams: a 3000-mutant ISM keeps only 56 of the 148
    # SMs busy inside the persistent recurrence kernel, whose time does not depend on its grid size
    if want('danq'):
        torch.manual_seed(0)
        azd2 = AnalyzeSequences(M.DanQ(1000, 919), sequence_length=1000, features=feats, reference_sequence=Genome)
        streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        azs = [azd, azd2]
        for k in range(4):
            with torch.cuda.stream(streams[k & 1]):
                azs[k & 1].ism_scores(seq, ("abs_diffs", "logits"), to_host=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(2 * n_seq):
            with torch.cuda.stream(streams[k & 1]):
                azs[k & 1].ism_scores(seq, ("abs_diffs", "logits"), to_host=False)
        torch.cuda.synchronize()
        dt2 = (time.perf_counter() - t0) / (2 * n_seq)
        emit(config="4: DanQ ISM, two sequences in flight on two streams, per 1000-bp sequence of 3000 mutants",
             precision="bf16", seconds_per_sequence=dt2, mutants_per_s=3000 / dt2,
             tflops=3000 / dt2 * azd.net.flops_per_seq / 1e12)

    # config 4 through the public API: ISM over a FASTA file of sequences, both TSV files per sequence written
    if want('danq'):
        import shutil, tempfile
        tmpd = tempfile.mkdtemp(prefix="sb_ismf_")
        nrec = 8 if args.quick else 24
        r11 = np.random.default_rng(11)
        with open(os.path.join(tmpd, "q.fa"), "w") as fh:
            for i in range(nrec):
                fh.write(">seq%d\n%s\n" % (i, "".join(np.array(list("ACGT"))[r11.integers(0, 4, 1000)].tolist())))
        azd.in_silico_mutagenesis_from_file(os.path.join(tmpd, "q.fa"), ["abs_diffs", "logits"], os.path.join(tmpd, "w"))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        azd.in_silico_mutagenesis_from_file(os.path.join(tmpd, "q.fa"), ["abs_diffs", "logits"], os.path.join(tmpd, "o"))
        torch.cuda.synchronize()
        dtf = (time.perf_counter() - t0) / nrec
        shutil.rmtree(tmpd, ignore_errors=True)
        emit(config="4: DanQ in_silico_mutagenesis_from_file (FASTA in, abs_diffs + logits TSV files out), per sequence",
             precision="bf16", seconds_per_sequence=dtf, mutants_per_s=3000 / dtf)

    # ---------------------------------------------------------------- config 5: Seqweaver / HeartENN sweep
    for name, model in ((("Seqweaver(217)", M.Seqweaver(217)), ("HeartENN(1000, 90)", M.HeartENN(1000, 90)),
                         ("DeeperDeepSEA(1000, 919) [SURVEY 8 f4]", M.DeeperDeepSEA(1000, 919)),
                         ("NonStrandSpecific(DeepSEA(1000, 919), mean) [SURVEY 8 f3]",
                          M.NonStrandSpecific(M.DeepSEA(1000, 919), "mean")))
                        if want('sweep') else ()):
        torch.manual_seed(0)
        model.eval()
        from selene_b200.nets import unwrap_non_strand_specific
        strand = unwrap_non_strand_specific(model)[1]
        net = B200Net(layers_from_module(model), 1000, strand_mode=strand)
        for batch in (64, 256, 1024, 4096):
            codes_b = torch.from_numpy(rng.integers(0, 4, (batch, 1000)).astype(np.uint8)).cuda()
            src = torch.arange(batch, dtype=torch.int32, device="cuda").repeat_interleave(2)
            p_ = torch.full((2 * batch,), 499, dtype=torch.int32, device="cuda")
            sc = torch.from_numpy(rng.integers(0, 4, 2 * batch).astype(np.uint8)).cuda()
            fn = lambda: net.forward_score(codes_b, src, p_, sc, 2 * batch, _lib.SB_PAIR_INTERLEAVED, precision="bf16",
                                           want=("abs_diffs", "logits"))
            t = timed(fn, warmup=2, iters=5)
            tf = 2 * batch * net.flops_per_seq * (2 if strand else 1) / t / 1e12   # both strands are computed
            emit(config="5: %s VEP-style scoring" % name, variants_per_batch=batch, ms=t * 1e3,
                 predictions_per_s=2 * batch / t, tflops=tf, tensor_frac_of_sustained=tf / float(peaks["bf16_tflops_sustained"]))

    # ---------------------------------------------------------------- config 3 (sampling half): sampler rate
    from selene_b200.samplers import RandomPositionsSampler
    import gzip
    bed = "/tmp/bench_feat.bed.gz"
    order = np.lexsort((s_i, c_i))
    with gzip.open(bed, "wt", compresslevel=1) as fh:
        for k in order[:300000]:
            fh.write("%s\t%d\t%d\tf%d\n" % (names[c_i[k]], s_i[k], s_i[k] + ln[k], f_i[k]))
    smp = RandomPositionsSampler(G, bed, feats, seed=436, validation_holdout=['chr6', 'chr7'],
                                 test_holdout=['chr8', 'chr9'], sequence_length=1000, center_bin_to_predict=200,
                                 feature_thresholds=0.5)
    if want('train'):
        # ------------------------------------------------------------ config 3: training step, batch 64
        from selene_b200.train_model import B200Trainer
        for prec in ("bf16", "fp32"):
            torch.manual_seed(0)
            mdl = M.DeepSEA(1000, 919)
            tr = B200Trainer(mdl, 1000, lr=0.01, momentum=0.9, weight_decay=1e-6,
                             dropout=B200Trainer.dropout_of_module(mdl), precision=prec)
            for batch in ((64, 256) if prec == "bf16" else (64,)):
                xb, yb = smp.sample_device(batch, seq_dtype="float32", label_dtype="float32")
                for _ in range(3):
                    tr.step(xb, yb)
                torch.cuda.synchronize()
                ns = 10 if prec == "bf16" else 4
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(ns):
                    tr.forward_backward(xb, yb)
                    tr.sgd_step()
                e1.record()
                torch.cuda.synchronize()
                dt_ = e0.elapsed_time(e1) / ns * 1e-3
                emit(config="3: DeepSEA training step (fwd train-mode + BCE + bwd + SGD), per GPU", precision=prec,
                     batch=batch, seconds_per_step=dt_, samples_per_s=batch / dt_,
                     tflops=batch * 3 * 1.0986e9 / dt_ / 1e12)
            if prec == "bf16":
                t0 = time.perf_counter()
                ns = 10
                for _ in range(ns):
                    xb, yb = smp.sample_device(64, seq_dtype="float32", label_dtype="float32")
                    tr.step(xb, yb)
                torch.cuda.synchronize()
                dt_ = (time.perf_counter() - t0) / ns
                # the same with the next batch drawn while the GPU steps (loss of every step still read on the host)
                t0 = time.perf_counter()
                nxt = smp.sample_device(64, seq_dtype="float32", label_dtype="float32")
                for _ in range(ns):
                    tr.step_async(*nxt)
                    nxt = smp.sample_device(64, seq_dtype="float32", label_dtype="float32")
                    tr.loss_value()
                torch.cuda.synchronize()
                dt_pipe = (time.perf_counter() - t0) / ns
                emit(config="3: sample + training step end to end, next batch drawn during the step", precision=prec, batch=64,
                     seconds_per_step=dt_pipe, samples_per_s=64 / dt_pipe)
                emit(config="3: sample + training step end to end (RandomPositionsSampler -> step -> loss to host)",
                     precision=prec, batch=64, seconds_per_step=dt_, samples_per_s=64 / dt_)
            del tr
    if not want('sampler'):
        return 0
    smp.sample_device(64)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    nb = 20
    for _ in range(nb):
        smp.sample_device(64)
    torch.cuda.synchronize()
    dt_ = (time.perf_counter() - t0) / nb
    emit(config="3 (sampling half): RandomPositionsSampler.sample_device(64), 919 features", seconds_per_batch=dt_,
         samples_per_s=64 / dt_)
    t0 = time.perf_counter()
    for _ in range(5):
        smp.sample(64)
    dt_ = (time.perf_counter() - t0) / 5
    emit(config="3 (sampling half): RandomPositionsSampler.sample(64) -> float64 numpy (reference signature)",
         seconds_per_batch=dt_, samples_per_s=64 / dt_)
    return 0


if __name__ == "__main__":
    sys.exit(main())
