"""Summarise an .ncu-rep (read on the CPU box): python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [batch] > profiles/kernels_rNN.md"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "smsp__inst_executed.sum", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
print("# ncu --set full summaries (%s, batch %d under ncu)\n" % (rep.split("/")[-1], batch))
print("```")
for r in data:
    name = r[col["Kernel Name"]]
    print("## %s  (launch id %s)" % (name[:90], r[col["ID"]]))
    for m in METRICS:
        if m in col:
            print("  %-62s %s %s" % (m, r[col[m]], units[col[m]]))
    pipes = {h.split("sm__inst_executed_pipe_")[1].split(".")[0]: r[i] for h, i in col.items()
             if h.startswith("sm__inst_executed_pipe_") and h.endswith("pct_of_peak_sustained_active")}
    if pipes:
        print("  pipes (% of peak, active): " + ", ".join("%s %s" % (k, v) for k, v in sorted(pipes.items()) if float(v or 0) >= 1.0))
    stalls = {h.split("smsp__average_warps_issue_stalled_")[1].split("_per_issue_active")[0]: float(r[i] or 0) for h, i in col.items()
              if h.startswith("smsp__average_warps_issue_stalled_") and "per_issue_active" in h and "not_issued" not in h}
    if stalls:
        tot = sum(stalls.values()) or 1.0
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:6]
        print("  stalls (share of stalled warp-cycles): " + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in top))
print("```")
# per-instruction view
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
print("\n## hottest SASS lines (warp-level samples)\n\n```")
blocks = src.split("\n\n")
cur = None
lines = list(csv.reader(io.StringIO(src)))
hidx = None
kern = "?"
acc = []
def flush():
    if not acc:
        return
    tot = sum(a[0] for a in acc) or 1
    ops = defaultdict(lambda: [0, 0])
    ninst = sum(a[1] for a in acc) or 1
    for s, n, txt in acc:
        op = txt.split()[0] if not txt.startswith("@") else txt.split()[1]
        op = op.split(".")[0]
        ops[op][0] += n
        ops[op][1] += s
    print("===== %s | SASS lines %d | warp inst %d | samples %d" % (kern[:80], len(acc), ninst, tot))
    for op, (n, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:14]:
        print("   %-10s %5.2f%% inst  %5.2f%% samples" % (op, 100.0 * n / ninst, 100.0 * s / tot))
    for s, n, txt in sorted(acc, key=lambda a: -a[0])[:12]:
        print("  %5.2f%%  x%-8d %s" % (100.0 * s / tot, n, txt[:70]))
for row in lines:
    if not row:
        continue
    if row[0].startswith("Kernel Name") or (len(row) == 1 and "(" in row[0]):
        continue
    if "Source" in row and "Sampling Data (All)" in " ".join(row) or (row and row[0] == "Address"):
        flush()
        acc = []
        hidx = {h: i for i, h in enumerate(row)}
        continue
    if hidx is None:
        continue
    try:
        txt = row[hidx["Source"]]
        s = int(float(row[hidx.get("Warp Stall Sampling (All Samples)", hidx.get("# Samples", 0))] or 0))
        n = int(float(row[hidx["Instructions Executed"]] or 0))
        acc.append((s, n, txt))
    except Exception:
        pass
flush()
print("```")
