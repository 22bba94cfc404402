import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as G
from oracle import synth, tif_oracle
pkg = G.load_package()
erp_h, wt, pad, off = int(sys.argv[1]), int(sys.argv[2]), 0.3, float(sys.argv[3])
src, tar = synth.synth_images(erp_h, 21)
flows = synth.synth_flows(wt, 20, off)
for blend in ("straightforward", "normwarp"):
  for wrap in (True, False):
    got = pkg.projection_icosahedron.ico2erp_flow(flows, erp_h, pad, src, tar, wrap, blend)
    want, dbg = tif_oracle.ico2erp_flow(flows, erp_h, pad, src, tar, wrap, blend, return_debug=True)
    err = np.abs(got - want).max(axis=2)
    exc = (dbg["seam_dist"] < 1e-5) | ((dbg["edge_dist"] < 1e-3) if blend == "normwarp" else False)
    if not wrap:
        exc = exc | (dbg["fu_spread"] > 0.25 * 2 * erp_h)
    bad = (err > 1e-3) & ~exc
    rr, cc = np.nonzero(bad)
    print(blend, "wrap", wrap, "bad", len(rr), "max err (unexcused)", (err * ~exc).max(), "p99.99", np.percentile(err, 99.99), "median", np.median(err))
    for r, c in list(zip(rr, cc))[:10]:
        print("  pix", r, c, "got", got[r, c], "want", want[r, c], "seam", dbg["seam_dist"][r, c], "edge", dbg["edge_dist"][r, c])
    if wrap:
        rows_bad = np.bincount(np.nonzero(bad)[0], minlength=erp_h)
        edges = [0, 1, 2, 4, 8, 16, 32, 64, 128, erp_h // 2]
        for a, b in zip(edges, edges[1:]):
            print("   rows [%d,%d) and mirror: bad %d" % (a, b, rows_bad[a:b].sum() + rows_bad[erp_h - b:erp_h - a].sum()))
