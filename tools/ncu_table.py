#!/usr/bin/env python
"""Per-kernel table from `ncu --metrics ... --csv` logs (long format): tools/ncu_table.py <csv>... ; averages over the launches."""
import collections
import csv
import re
import sys

rows = []
for path in sys.argv[1:]:
    rows += list(csv.DictReader(l for l in open(path) if l.startswith('"')))
per = collections.OrderedDict()
for r in rows:
    name = re.sub(r"<unnamed>::|\(.*", "", r["Kernel Name"]).split("<")[0]
    k = per.setdefault(name, collections.defaultdict(list))
    try:
        k[r["Metric Name"]].append(float(r["Metric Value"].replace(",", "")))
    except ValueError:
        pass
    k["_unit_" + r["Metric Name"]] = r["Metric Unit"]


def avg(k, m):
    v = k.get(m)
    return sum(v) / len(v) if v else float("nan")


def to(k, m, scale):   # ncu scales units per row: normalise
    u = k.get("_unit_" + m, "")
    f = {"nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
         "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    return avg(k, m) * f / scale


print("%-28s %4s %10s %10s %10s %9s %7s %7s %5s %6s %5s" % ("kernel", "n", "time_us", "read_MB", "write_MB", "GB/s", "dram%", "issue%", "regs", "grid", "block"))
for name, k in per.items():
    n = len(k["gpu__time_duration.sum"])
    t = to(k, "gpu__time_duration.sum", 1e-6)
    rd, wr = to(k, "dram__bytes_read.sum", 1e6), to(k, "dram__bytes_write.sum", 1e6)
    print("%-28s %4d %10.1f %10.2f %10.2f %9.1f %7.1f %7.1f %5d %6d %5d" % (
        name, n, t, rd, wr, (rd + wr) / t * 1e3 if t else 0.0, avg(k, "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
        avg(k, "smsp__issue_active.avg.pct_of_peak_sustained_active"), avg(k, "launch__registers_per_thread"), avg(k, "launch__grid_size"),
        avg(k, "launch__block_size")))
