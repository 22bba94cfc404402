"""GPU vs oracle over the full field at one ERP height; prints the worst unexcused pixels with their per-face terms."""
import sys
import os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
from oracle import synth, tif_oracle

erp_h = int(sys.argv[1]) if len(sys.argv) > 1 else 1920
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
entry.build()
pkg = entry.load_package()
src, tar, flows = synth.synth(erp_h, seed, 480)
for blend in ("straightforward", "normwarp"):
    want, dbg = tif_oracle.ico2erp_flow(flows, erp_h, 0.3, src, tar, True, blend, return_debug=True)
    got = pkg.projection_icosahedron.ico2erp_flow(flows, erp_h, 0.3, src, tar, True, blend)
    excuse = dbg["seam_dist"] < 1e-5
    if blend == "normwarp":
        excuse |= dbg["edge_dist"] < 1e-3
    err = np.abs(got - want).max(axis=2)
    bad = (err > 1e-3) & ~excuse
    print(blend, "max err", err.max(), "bad", int(bad.sum()), "median", np.median(err), "p99.99", np.quantile(err, 0.9999))
    order = np.argsort(-(err * ~excuse).ravel())[:12]
    for o in order:
        r, c = divmod(int(o), 2 * erp_h)
        print("  r %d c %d err %.3e got %s want %s spread %.3f seam %.2e edge %.2e" % (
            r, c, err[r, c], got[r, c], want[r, c], dbg["fu_spread"][r, c], dbg["seam_dist"][r, c], dbg["edge_dist"][r, c]))
