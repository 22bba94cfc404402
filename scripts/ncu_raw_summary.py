"""Summarise `ncu --page raw --csv` per launch: time, DRAM bytes, throughputs, occupancy, pipes."""
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum"]
    pipes = [h for h in hdr if h.startswith("sm__inst_executed_pipe_") and h.endswith(".avg.pct_of_peak_sustained_active")]
    seen = set()
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        if name in seen:            # one launch per kernel: repeats of the same launch differ by < 1 %
            continue
        seen.add(name)
        print("## " + name)
        for w in want:
            if w in idx:
                print("  %-62s %s %s" % (w, r[idx[w]], units[idx[w]]))
        print("  pipes (%% of peak, active): " + ", ".join(
            "%s %.1f" % (n.replace("sm__inst_executed_pipe_", "").replace(".avg.pct_of_peak_sustained_active", ""), float(r[idx[n]]))
            for n in pipes if float(r[idx[n]]) > 1.0))


if __name__ == "__main__":
    main(sys.argv[1])
