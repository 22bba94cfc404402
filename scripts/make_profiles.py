"""Turn the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

  python scripts/make_profiles.py <tag> <launches.csv> <full.ncu-rep> <batch used under ncu> <erp_h>
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import ncu_raw_summary  # noqa: E402
import ncu_source_summary  # noqa: E402

# substring of the ncu kernel name -> the kernel's name in bench.py's roofline.kernels
KERNEL_KEYS = [("k_erp2ico", "k_erp2ico"), ("k_lift_nw", "k_lift_nw"), ("k_blend_nw", "k_blend_nw"),
               ("k_lift<0", "k_lift<straightforward>"), ("k_lift<(int)0", "k_lift<straightforward>"),
               ("k_gather<0", "k_gather<straightforward>"), ("k_gather<(int)0", "k_gather<straightforward>"),
               ("k_warp_rgb_fast", "k_warp_backward")]


def capture(fn, *args):
    buf, old = io.StringIO(), sys.stdout
    sys.stdout = buf
    try:
        fn(*args)
    finally:
        sys.stdout = old
    return buf.getvalue()


def main():
    tag, launches, rep, batch, erp_h = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
    out = os.path.join(ROOT, "profiles")
    os.makedirs(out, exist_ok=True)
    # 1) launch list: per-kernel totals and shares (cold-cache, serialised: compare SHARES)
    rows = [r for r in csv.reader(open(launches)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows[1:]:
        if len(r) != len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[idx["Kernel Name"]].split("(")[0]
        val = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]]
        val *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        per.setdefault(name, []).append(val)
    total = sum(sum(v) for v in per.values())
    with open(os.path.join(out, "launches_%s.md" % tag), "w") as f:
        f.write("# ncu launch list (%s)\n\n`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` over `bench.py --steps 2 "
                "--warmup 1 --batch %d --no-e2e --no-cpu-baseline` (scripts/gpu_record.sh; the plan build and bench.py's "
                "per-kernel timing loops are part of the list).\nPer-launch times are cold-cache and "
                "serialised: compare shares, not absolutes.  Raw list: `launches_%s.csv`.\n\n" % (tag, batch, tag))
        f.write("| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|\n")
        for name, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            f.write("| `%s` | %d | %.1f | %.1f | %.1f %% |\n" % (name, len(v), sum(v), sum(v) / len(v), 100.0 * sum(v) / total))
    subprocess.run(["cp", launches, os.path.join(out, "launches_%s.csv" % tag)], check=True)
    # 2) full capture: raw + source summaries
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    open("/tmp/_raw.csv", "w").write(raw)
    open("/tmp/_src.csv", "w").write(src)
    with open(os.path.join(out, "kernels_%s.md" % tag), "w") as f:
        f.write("# ncu --set full summaries (%s, batch %d under ncu, ERP %dx%d)\n\n```\n" % (tag, batch, 2 * erp_h, erp_h))
        f.write(capture(ncu_raw_summary.main, "/tmp/_raw.csv"))
        f.write("```\n\n## per-instruction view (opcode mix, stall reasons, hottest SASS lines)\n\n```\n")
        f.write(capture(ncu_source_summary.main, "/tmp/_src.csv", 10))
        f.write("```\n")
    # 3) DRAM traffic per frame pair for bench.py's roofline.traffic
    rr = list(csv.reader(io.StringIO(raw)))
    h, units = rr[0], rr[1]
    ix = {n: i for i, n in enumerate(h)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    traffic = {}
    for r in rr[2:]:
        name = r[ix["Kernel Name"]]
        tot = sum(float(r[ix[m]]) * scale.get(units[ix[m]], 1.0) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        for key, label in KERNEL_KEYS:
            if key in name:
                traffic.setdefault(label, []).append(tot / batch)
    traffic = {k: max(v) for k, v in traffic.items()}
    # bench.py looks the figures up per configuration ("h<ERP height>_l<subdivision level>"); other configurations of
    # an earlier capture are kept
    path = os.path.join(out, "traffic.json")
    try:
        doc = json.load(open(path))
        if not isinstance(doc.get("bytes_per_pair", {}).get("h%d_l0" % erp_h, {}), dict) or "tag" in doc:
            doc = {"bytes_per_pair": {k: v for k, v in doc.get("bytes_per_pair", {}).items() if isinstance(v, dict)}}
    except Exception:
        doc = {"bytes_per_pair": {}}
    doc["bytes_per_pair"]["h%d_l0" % erp_h] = traffic
    doc.setdefault("captures", {})["h%d_l0" % erp_h] = {"tag": tag, "batch_under_ncu": batch}
    doc["note"] = ("dram__bytes_read.sum + dram__bytes_write.sum of one launch / batch under ncu --set full; k_erp2ico and "
                   "k_warp_backward are per image, the stitch kernels per frame pair")
    json.dump(doc, open(path, "w"), indent=1)
    print(open(os.path.join(out, "launches_%s.md" % tag)).read())
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
