"""Summarise an ncu --metrics gpu__time_duration.sum --csv launch list: per-kernel time, share and launch count."""
import collections, csv, re, sys

def summarize(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"^.*::", "", name)
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    out = [f"total {T:.1f} us over {sum(cnt.values())} launches"]
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:top]:
        out.append(f"{v:10.1f} us {100 * v / T:5.1f}%  n={cnt[k]:4d}  {k}")
    return "\n".join(out)

if __name__ == "__main__":
    print(summarize(sys.argv[1]))
