#!/usr/bin/env python
"""Per-source-line stall samples / instructions from `ncu -i X --page source --csv --print-source cuda,sass --kernel-name ...`:
tools/ncu_source_lines.py file.csv [top_n]"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
lines = []
cur_file = ""
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1]
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r or not r[0].strip().isdigit():
        continue
    try:
        samples = int(r[4]) if r[4] not in ("-", "") else 0
        inst = int(r[7]) if r[7] not in ("-", "") else 0
    except ValueError:
        continue
    lines.append((samples, inst, cur_file.split("/")[-1], int(r[0]), r[1]))
tot_s = sum(x[0] for x in lines) or 1
tot_i = sum(x[1] for x in lines) or 1
print("total samples %d, total warp instructions %d" % (tot_s, tot_i))
for s, i, f, ln, src in sorted(lines, reverse=True)[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%-4d %s" % (100.0 * s / tot_s, 100.0 * i / tot_i, f, ln, src.strip()[:110]))
