"""Text summary of an `ncu --set full` capture for profiles/: python tools/ncu_summary.py file.ncu-rep [out.txt]
          [--traffic-key B]   (also records dram read / write bytes of the first kernel in profiles/roofline_traffic.json)"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "sm__sass_inst_executed_op_shared_ld.sum"]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    rep = sys.argv[1]
    args = [a for a in sys.argv[2:] if not a.startswith("--")]
    out = args[0] if args else None
    key = sys.argv[sys.argv.index("--traffic-key") + 1] if "--traffic-key" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f"# {os.path.basename(rep)}: ncu --set full --clock-control none (values per launch)"]
    traffic = None
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        lines.append(f"kernel: {d.get('Kernel Name', '?')[:120]}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}")
        for k in WANT:
            if k in d:
                lines.append(f"  {k:70s} {d[k]:>16s} {u.get(k, '')}")
        if traffic is None and "dram__bytes_read.sum" in d:
            traffic = (to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]), to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"]))
    text = "\n".join(lines) + "\n"
    if out:
        with open(out, "w") as f:
            f.write(text)
    print(text)
    if key and traffic:
        path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        rec = json.load(open(path)) if os.path.exists(path) else {}
        rec[str(key)] = {"dram_read_bytes": traffic[0], "dram_write_bytes": traffic[1], "source": os.path.relpath(out or rep, ROOT)}
        json.dump(rec, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
