#!/usr/bin/env python
"""Per-source-line stall samples / warp instructions from
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K > K.csv
(one CSV row per SASS instruction, tagged with its source line): tools/ncu_source_lines.py K.csv [top_n]"""
import csv
import sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr = "", None
agg = defaultdict(lambda: [0, 0, 0, ""])       # (file, line) -> samples, warp inst, thread inst, source
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        i_smp, i_inst, i_thr = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or len(r) <= i_thr or not r[0].strip().isdigit():
        continue

    def num(x):
        try:
            return int(float(x))
        except ValueError:
            return 0
    a = agg[(cur_file, int(r[0]))]
    a[0] += num(r[i_smp]); a[1] += num(r[i_inst]); a[2] += num(r[i_thr]); a[3] = r[1]
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
tot_t = sum(a[2] for a in agg.values()) or 1
print("total samples %d, warp instructions %d, thread instructions %d (%.1f lanes/inst)" % (tot_s, tot_i, tot_t, tot_t / tot_i))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% smp %5.1f%% inst %4.1f lanes  %s:%-4d %s" % (100.0 * a[0] / tot_s, 100.0 * a[1] / tot_i, a[2] / max(a[1], 1), f, ln, a[3].strip()[:105]))
