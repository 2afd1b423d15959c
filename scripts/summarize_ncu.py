"""Summarises gpurun_out/launches.csv (ncu launch list) and a .ncu-rep (ncu --set full) into a
markdown file under profiles/.  Usage: python scripts/summarize_ncu.py <tag> [rep ...]"""
import collections
import csv
import subprocess
import sys

tag = sys.argv[1]
reps = sys.argv[2:]
out = open("profiles/%s_summary.md" % tag, "w")
rows = list(csv.reader(open("gpurun_out/launches.csv")))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[hi + 2:]:
    if len(r) <= vi:
        continue
    a = agg.setdefault(r[ki][:70], [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(',', ''))
tot = sum(a[1] for a in agg.values())
out.write("# ncu launch list (%s)\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over "
          "`python bench.py --steps 2 --warmup 1 --no-cpu --no-hbm --no-train --strong-variants 0` (the bench step at its real size: 16 384 variants, 240 Mbp genome) "
          "(cold-cache, serialised: compare shares)\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n" % tag)
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.write("| `%s` | %d | %.1f | %.1f%% |\n" % (k, a[0], a[1] / 1e3, 100 * a[1] / tot))
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'launch__grid_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__cycles_active.avg']
for rep in reps:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(txt.splitlines()))
    hh = rr[0]
    out.write("\n# ncu --set full: %s\n\n" % rep)
    for r in rr[2:]:
        out.write("\n| metric | value | unit |\n|---|---|---|\n")
        for w in want:
            for i, x in enumerate(hh):
                if x == w or x.startswith(w):
                    out.write("| %s | %s | %s |\n" % (x, r[i], rr[1][i]))
                    break
out.close()
print(open("profiles/%s_summary.md" % tag).read())
