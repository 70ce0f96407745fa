"""Print selected metrics of an .ncu-rep: python tools/ncu_metrics.py file.ncu-rep [substring ...]"""
import csv, subprocess, sys
rep = sys.argv[1]
want = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
                        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct",
                        "launch__registers_per_thread", "sm__throughput.avg.pct", "l1tex__throughput.avg.pct", "lts__throughput.avg.pct",
                        "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
                        "smsp__average_warp", "smsp__average_warps_issue_stalled", "launch__grid_size", "launch__waves"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
for r in rows[2:]:
    print("==", r[h.index("Kernel Name")] if "Kernel Name" in h else "")
    for i, k in enumerate(h):
        if any(w in k for w in want):
            print(f"  {k} = {r[i]} {rows[1][i]}")
