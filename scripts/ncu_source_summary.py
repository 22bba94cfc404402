"""Summarise `ncu --page source --csv` output: opcode mix, stall reasons, hottest SASS lines per kernel."""
import collections
import csv
import sys


def main(path, top=12):
    rows = list(csv.reader(open(path)))
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            kernels.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["data"].append(r)
    seen = set()
    for k in kernels:
        if k["name"] in seen:       # one launch per kernel
            continue
        seen.add(k["name"])
        hdr, data = k["hdr"], k["data"]
        idx = {h: i for i, h in enumerate(hdr)}
        ie = idx["Instructions Executed"]
        tot = sum(int(r[ie]) for r in data)
        nsamp = sum(int(r[idx["# Samples"]]) for r in data)
        print("=====", k["name"][:70], "| SASS lines", len(data), "| warp inst", tot, "| samples", nsamp)
        byop, samp = collections.Counter(), collections.Counter()
        for r in data:
            toks = r[idx["Source"]].split()
            op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
            byop[op] += int(r[ie])
            samp[op] += int(r[idx["# Samples"]])
        for op, n in byop.most_common(top * 2):
            print("   %-10s %6.2f%% inst  %6.2f%% samples" % (op, 100.0 * n / tot, 100.0 * samp[op] / max(nsamp, 1)))
        st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tots = {h: sum(int(r[idx[h]] or 0) for r in data) for h in st}
        s = sum(tots.values()) or 1
        print("   stalls:", {h: round(100.0 * v / s, 1) for h, v in sorted(tots.items(), key=lambda kv: -kv[1])[:8]})
        hot = sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:top]
        for r in hot:
            print("   %5.2f%%  x%-9s %s" % (100.0 * int(r[idx["# Samples"]]) / max(nsamp, 1), r[ie], r[idx["Source"]].strip()[:90]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 12)
