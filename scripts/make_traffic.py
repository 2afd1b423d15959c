"""profiles/traffic.json from an `ncu --set full` capture of the bench step: DRAM bytes (read + write) of one launch of
the dominant kernel (the longest tc_layer_kernel launch = conv2 of a full chunk).
    python scripts/make_traffic.py gpurun_out/prof_tc.ncu-rep <rows per launch> <tag>"""
import csv, json, subprocess, sys
rep, rows, tag = sys.argv[1], int(sys.argv[2]), sys.argv[3]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(txt.splitlines()))
h, u = rr[0], rr[1]
ik, it, ir, iw = h.index("Kernel Name"), h.index("gpu__time_duration.sum"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
best = None
for r in rr[2:]:
    if "tc_layer_kernel" not in r[ik]:
        continue
    t = float(r[it].replace(",", ""))
    if best is None or t > best[0]:
        best = (t, float(r[ir].replace(",", "")) * scale[u[ir]], float(r[iw].replace(",", "")) * scale[u[iw]], u[it])
out = {"conv2": {"rows_per_launch": rows, "dram_bytes": best[1] + best[2], "dram_bytes_read": best[1], "dram_bytes_write": best[2],
                 "duration": "%g %s" % (best[0], best[3]),
                 "source": "profiles/%s_summary.md (ncu --set full --clock-control none, longest tc_layer_kernel launch)" % tag}}
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print(json.dumps(out))
