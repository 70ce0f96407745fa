"""Shared-memory wavefronts per source line of an ncu --set full --import-source capture: python tools/ncu_smem.py file.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur, hdr, recs = None, None, []
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0]: continue
    d = {}
    for k, v in zip(hdr, r): d.setdefault(k, v)
    try:
        w = int(d.get("L1 Wavefronts Shared", "0") or 0); ideal = int(d.get("L1 Wavefronts Shared Ideal", "0") or 0)
    except ValueError:
        continue
    if w: recs.append((w, ideal, cur, d["Line No"], d["Source"].strip()[:100]))
tot = sum(x[0] for x in recs); toti = sum(x[1] for x in recs)
print(f"shared wavefronts {tot} (ideal {toti})")
for w, i, f, ln, src in sorted(recs, key=lambda x: -x[0])[:top]:
    print(f"{100*w/tot:5.1f}%  {w:>10d} (ideal {i:>10d})  {f}:{ln:>4}  {src}")
