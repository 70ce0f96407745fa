"""Per-source-line summary of an ncu --set full --import-source capture:
python tools/ncu_source.py file.ncu-rep [top N]   -> lines ranked by warp-stall samples, with executed instructions."""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, recs = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0]:
        continue          # SASS rows have an empty line number
    d = {}
    for k, v in zip(hdr, r):
        d.setdefault(k, v)
    try:
        samples = int(d.get("# Samples", "0") or 0)
        inst = int(d.get("Instructions Executed", "0") or 0)
    except ValueError:
        continue
    recs.append((samples, inst, cur, d["Line No"], d["Source"].strip()[:100], d))
tot_s = sum(x[0] for x in recs) or 1
tot_i = sum(x[1] for x in recs) or 1
print(f"total samples {tot_s}, total warp instructions {tot_i}")
for s, i, f, ln, src, d in sorted(recs, key=lambda x: -x[0])[:top]:
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v and v.isdigit() and int(v) > 0}
    top3 = ", ".join(f"{k}:{v}" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:3])
    bc = d.get("L1 Wavefronts Shared Excessive", "")
    print(f"{100*s/tot_s:5.1f}% smp {100*i/tot_i:5.1f}% inst  {f}:{ln:>4}  {src}   [{top3}] exc_wf={bc}")
