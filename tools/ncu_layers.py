#!/usr/bin/env python
"""Turn an ncu per-launch CSV of one layer-by-layer pass into a table: duration, DRAM bytes, DRAM % of peak, tensor-pipe %.
Capture (on a B200 box, after the same command ran cleanly without ncu):
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \\
      --clock-control none --csv --log-file out.csv python tools/run_once.py 8192 layered 1
usage: ncu_layers.py out.csv [title] > table.md"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
iI, iK, iM, iU, iV = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
launches = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= iV:
        continue
    d = launches.setdefault(r[iI], {"kernel": r[iK]})
    v = float(r[iV].replace(",", ""))
    u = r[iU]
    if r[iM] == "gpu__time_duration.sum":
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v          # -> us
    if r[iM].startswith("dram__bytes"):
        v = v / 1e6 if u == "byte" else v / 1e3 if u == "Kbyte" else v * 1e3 if u == "Gbyte" else v   # -> MB
    d[r[iM]] = v
print("# %s" % (sys.argv[2] if len(sys.argv) > 2 else "ncu per-launch metrics"))
print()
print("| # | kernel | us | DRAM read MB | DRAM write MB | DRAM % of peak | tensor pipe % |")
print("|---|---|---|---|---|---|---|")
tot = 0.0
for i, d in enumerate(launches.values()):
    name = d["kernel"].replace("yf::", "").replace("void ", "")[:44]
    t = d.get("gpu__time_duration.sum", 0.0); tot += t
    print("| %d | %s | %.1f | %.1f | %.1f | %.1f | %.2f |" % (i, name, t, d.get("dram__bytes_read.sum", 0), d.get("dram__bytes_write.sum", 0),
          d.get("dram__throughput.avg.pct_of_peak_sustained_elapsed", 0), d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0)))
print()
print("sum of launch durations: %.1f us" % tot)
