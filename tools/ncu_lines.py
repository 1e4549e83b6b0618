#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by source line.
usage: ncu_lines.py dump.csv kernel_substring [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
def num(x):
    try: return int(x.replace(',', ''))
    except Exception: return 0
allrows = []
i = 0
fname = func = None
while i < len(rows):
    r = rows[i]
    if r and r[0] == "File Path": fname = r[1].split('/')[-1]
    elif r and r[0] == "Function Name": func = r[1]
    elif r and r[0] == "Line No":
        hdr = r; col = {h: k for k, h in enumerate(hdr)}
        i += 1
        while i < len(rows) and not (rows[i] and rows[i][0] in ("File Path", "Function Name", "Line No")):
            q = rows[i]; i += 1
            if len(q) < 8 or not q[0].isdigit() or want not in (func or ""): continue
            allrows.append((fname, int(q[0]), q[1], num(q[col["# Samples"]]), num(q[col["Instructions Executed"]])))
        continue
    i += 1
tot = sum(r[3] for r in allrows) or 1
print("kernel ~", want, "total samples", tot, "total warp instructions", sum(r[4] for r in allrows))
for f, ln, src, smp, inst in sorted(allrows, key=lambda r: -r[3])[:top]:
    print("%-14s %5d %7d %5.1f%% inst=%9d  %s" % (f, ln, smp, 100 * smp / tot, inst, src.strip()[:100]))
