"""SASS evidence for profiles/: per hot kernel of libtif_b200.so the instruction count, the opcode mix, the marker
mnemonics (LDGSTS = cp.async, UTMALDG = TMA tensor load, SYNCS = mbarrier, REDUX, MUFU, ATOMS ...) and the body of its
largest loop.   python scripts/sass_excerpt.py > profiles/sass_r02.md     (runs on the CPU box: cuobjdump + nvdisasm)"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "360opticalflow-tangentimages_b200", "libtif_b200.so")
HOT = ["k_erp2ico_staged", "k_erp2ico_tma", "k_lift_nw", "k_blend_nw", "k_liftILi0E", "k_gatherILi0E", "k_warp_rgb_fastIf"]
MARKERS = ["LDGSTS", "LDGDEPBAR", "UTMALDG", "UBLKCP", "SYNCS", "REDUX", "MUFU", "ATOMS", "ATOMG", "RED", "SHFL", "VOTE", "BAR", "STG", "LDG", "LDS",
           "STS", "DFMA", "DADD", "DMUL", "HFMA2", "I2F", "F2I", "FRND"]


def main():
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, capture_output=True)
    print("# SASS of the hot kernels (libtif_b200.so, sm_100a, `cuobjdump -xelf` + `nvdisasm -c`)\n")
    for cubin in sorted(os.listdir(tmp)):
        if not cubin.endswith(".cubin") or "tif_" not in cubin:
            continue
        txt = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
        funcs, cur = collections.OrderedDict(), None
        for line in txt.splitlines():
            m = re.match(r"^\.text\.(\S+):", line)
            if m:
                cur = m.group(1)
                funcs[cur] = []
            elif cur is not None:
                funcs[cur].append(line)
        for name, lines in funcs.items():
            if not any(h in name for h in HOT) or "$" in name:
                continue
            inst = []           # (label or None, text)
            for ln in lines:
                m = re.match(r"^\s*/\*[0-9a-f]{4,}\*/\s+(.*?)\s*;", ln)
                if m:
                    inst.append((None, m.group(1)))
                else:
                    m = re.match(r"^(\.L_x_\d+):", ln)
                    if m:
                        inst.append((m.group(1), None))
            ops = collections.Counter()
            for _, t in inst:
                if t is None:
                    continue
                toks = t.split()
                op = toks[1] if toks[0].startswith("@") else toks[0]
                ops[op.split(".")[0]] += 1
            total = sum(ops.values())
            demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
            print("## `%s`\n" % demangled)
            print("%d SASS instructions.  Opcode mix: %s\n" % (total, ", ".join("%s %d" % kv for kv in ops.most_common(14))))
            print("Markers: " + ", ".join("%s %d" % (k, ops[k]) for k in MARKERS if ops.get(k)) + "\n")
            # largest loop: a backward branch to a label seen earlier
            pos, label_at = 0, {}
            seq = []
            for lab, t in inst:
                if lab:
                    label_at[lab] = len(seq)
                else:
                    seq.append(t)
            best = None
            for i, t in enumerate(seq):
                m = re.search(r"BRA\S*\s+.*`\((\.L_x_\d+)\)", t)
                if m and m.group(1) in label_at and label_at[m.group(1)] <= i:
                    n = i - label_at[m.group(1)] + 1
                    if best is None or n > best[1]:
                        best = (label_at[m.group(1)], n)
            if best:
                a, n = best
                print("Largest loop: %d instructions; its first %d:\n\n```" % (n, min(n, 70)))
                for t in seq[a:a + min(n, 70)]:
                    print("    " + t)
                print("```\n")


if __name__ == "__main__":
    main()
