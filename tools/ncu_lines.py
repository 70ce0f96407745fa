"""Instruction / stall-sample share per source line, in file order: python tools/ncu_lines.py file.ncu-rep [min_pct]"""
import csv, subprocess, sys
rep = sys.argv[1]
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, recs = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0]:
        continue
    d = {}
    for k, v in zip(hdr, r):
        d.setdefault(k, v)
    try:
        s = int(d.get("# Samples", "0") or 0); i = int(d.get("Instructions Executed", "0") or 0)
    except ValueError:
        continue
    recs.append((cur, int(d["Line No"]), s, i, d["Source"].strip()[:90]))
ts = sum(x[2] for x in recs) or 1; ti = sum(x[3] for x in recs) or 1
for f, ln, s, i, src in recs:
    if 100 * s / ts >= minp or 100 * i / ti >= minp:
        print(f"{f}:{ln:>4} {100*s/ts:5.1f}% smp {100*i/ti:5.1f}% inst  {src}")
