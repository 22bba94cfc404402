"""Per CUDA source line: warp instructions and stall samples of one kernel.
usage: python scripts/ncu_lines.py report.ncu-rep kernel_substring [min_pct]"""
import csv, io, subprocess, sys, collections
rep, want = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
agg = collections.OrderedDict()
fpath, fn, hdr, done = None, None, None, set()
active = False
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        fn = r[1]; active = want in fn; continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if not active or hdr is None: continue
    if r[0] and r[0].isdigit():
        key = (fpath, int(r[0]))
        try:
            n = int(r[hdr["Instructions Executed"]]); s = int(r[hdr["# Samples"]])
        except Exception: continue
        a = agg.setdefault(key, [0, 0, r[1].strip()])
        a[0] += n; a[1] += s
tot = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
print("kernel ~ %s: warp inst %d, samples %d" % (want, tot, ts))
for (f, ln), (n, s, src) in sorted(agg.items()):
    if 100.0 * n / tot >= minpct or 100.0 * s / ts >= minpct:
        print("%-18s %5d  %5.2f%% inst %5.2f%% smp  %s" % (f, ln, 100.0 * n / tot, 100.0 * s / ts, src[:100]))
