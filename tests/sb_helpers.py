"""Synthetic inputs shared by the tests, bench.py and smoke() (seeded, small)."""
import os

import numpy as np


def synth_genome(n_chrom=3, length=20000, seed=1337, n_frac=0.01, lower_frac=0.05, odd_chars=True):
    """dict chrom -> str with N runs, soft-masked lowercase runs and a few odd characters."""
    rng = np.random.default_rng(seed)
    out = {}
    for c in range(n_chrom):
        L = length + 137 * c
        arr = np.array(list("ACGT"))[rng.integers(0, 4, L)]
        # telomere-like N blocks
        arr[:50] = "N"
        arr[-30:] = "N"
        for _ in range(max(1, int(n_frac * L / 40))):
            s = rng.integers(0, L - 60)
            arr[s:s + rng.integers(1, 60)] = "N"
        for _ in range(max(1, int(lower_frac * L / 100))):
            s = rng.integers(0, L - 200)
            e = s + rng.integers(1, 200)
            arr[s:e] = np.char.lower(arr[s:e])
        if odd_chars:
            for ch in "nRYU":
                arr[rng.integers(0, L, 5)] = ch
        out["chr%d" % (c + 1)] = "".join(arr.tolist())
    return out


def write_fasta(genome, path, width=60, with_fai=True):
    fai = []
    with open(path, "wb") as fh:
        for name, seq in genome.items():
            fh.write((">%s\n" % name).encode())
            off = fh.tell()
            for i in range(0, len(seq), width):
                fh.write(seq[i:i + width].encode() + b"\n")
            fai.append("%s\t%d\t%d\t%d\t%d\n" % (name, len(seq), off, width, width + 1))
    if with_fai:
        with open(path + ".fai", "w") as fh:
            fh.writelines(fai)


def synth_intervals(genome, n_features=12, per_feature=200, seed=7):
    """list of (chrom, start, end, feature) sorted by (chrom, start); overlapping, nested and
    abutting intervals included."""
    rng = np.random.default_rng(seed)
    rows = []
    chroms = list(genome.keys())
    for f in range(n_features):
        for _ in range(per_feature):
            c = chroms[rng.integers(0, len(chroms))]
            L = len(genome[c])
            ln = int(np.clip(rng.lognormal(np.log(300), 0.8), 5, 3000))
            s = int(rng.integers(0, max(1, L - ln)))
            rows.append((c, s, s + ln, "f%d" % f))
            if rng.random() < 0.3:  # overlapping / nested partner
                s2 = s + int(rng.integers(0, ln))
                rows.append((c, s2, s2 + int(rng.integers(1, ln + 50)), "f%d" % f))
    rows.sort(key=lambda r: (r[0], r[1], r[2]))
    return rows


def write_bed(rows, path):
    import gzip
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "wt") as fh:
        for c, s, e, f in rows:
            fh.write("%s\t%d\t%d\t%s\n" % (c, s, e, f))


def rows_by_chrom(rows):
    by = {}
    for c, s, e, f in rows:
        by.setdefault(c, []).append([c, str(s), str(e), f])
    return by


def stack_forward_rounded(x_bl4, layers, presigmoid=False):
    """The oracle's conv-stack forward (oracle.stack_forward) with the ROUNDING POINTS of the bf16 throughput
    path made explicit: weights rounded to bf16, every layer's (pooled) output rounded to bf16, fp32
    accumulation, biases and the last layer's output in fp32.  A checker for the tcgen05 path that is ~10x
    tighter than the fp32 oracle (what is left is summation order)."""
    import torch
    import torch.nn.functional as F
    bf = lambda t: t.to(torch.bfloat16).to(torch.float32)
    with torch.no_grad():
        x = torch.tensor(np.asarray(x_bl4, dtype=np.float32)).transpose(1, 2)
        flat = False
        for i, lay in enumerate(layers):
            w = bf(torch.tensor(np.asarray(lay["weight"], dtype=np.float32)))
            b = torch.tensor(np.asarray(lay["bias"], dtype=np.float32))
            if lay["type"] == "conv":
                x = F.conv1d(x, w, b)
                if lay.get("relu"):
                    x = F.relu(x)
                if lay.get("pool", 1) > 1:
                    x = F.max_pool1d(x, lay["pool"], lay["pool"])
            else:
                if not flat:
                    x = x.reshape(x.size(0), -1)
                    flat = True
                x = F.linear(x, w, b)
                if lay.get("relu"):
                    x = F.relu(x)
            if i < len(layers) - 1:
                x = bf(x)
        return (x if presigmoid else torch.sigmoid(x)).numpy()


def scaled_deepsea(scale, seed=0):
    """tests/sb_models.DeepSEA under torch.manual_seed(seed) with the three conv weights multiplied by ``scale``
    — the same weights as oracle.deepsea_layers(1000, 919, seed, scale)."""
    import torch
    import sb_models
    torch.manual_seed(seed)
    m = sb_models.DeepSEA(1000, 919)
    with torch.no_grad():
        for mod in m.conv_net:
            if isinstance(mod, torch.nn.Conv1d):
                mod.weight.mul_(scale)
    return m
