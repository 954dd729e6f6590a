#!/usr/bin/env python
"""Annotated SASS dump of one ncu report: tools/ncu_sass.py <report.ncu-rep> <out.txt>; prints hot branches."""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[1]
ia, isrc, iex, ith, ismp = (hdr.index(k) for k in ("Address", "Source", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
body = rows[2:]
tot = sum(int(r[iex]) for r in body); tots = sum(int(r[ismp]) for r in body)
base = int(body[0][ia], 16)
with open(out, "w") as f:
    f.write("# total warp instructions %d, samples %d\n" % (tot, tots))
    for n, r in enumerate(body):
        ex, th = int(r[iex]), int(r[ith])
        f.write("%4d %05x %6.3f%% ex=%11d thr=%4.1f smp=%5.2f%%  %s\n" % (n, int(r[ia], 16) - base, 100 * ex / tot, ex, th / ex if ex else 0, 100 * int(r[ismp]) / tots, r[isrc].strip()))
print("total", tot)
