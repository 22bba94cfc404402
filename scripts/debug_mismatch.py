import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as G
from oracle import synth, tif_oracle
pkg = G.load_package()
name, blend, wrap = sys.argv[1], sys.argv[2], bool(int(sys.argv[3]))
g = np.load("tests/golden/%s.npz" % name)
src, tar = synth.synth_images(int(g["erp_h"]), int(g["seed"]), bool(g["smooth"]))
flows = synth.synth_flows(int(g["tangent_w"]), 20, float(g["u_offset"]))
got = pkg.projection_icosahedron.ico2erp_flow(flows, int(g["erp_h"]), float(g["pad"]), src, tar, wrap, blend)
want, dbg = tif_oracle.ico2erp_flow(flows, int(g["erp_h"]), float(g["pad"]), src, tar, wrap, blend, return_debug=True)
err = np.abs(got - want).max(axis=2)
rr, cc = np.nonzero(err > 1e-3)
print("bad", len(rr), "max", err.max())
for r, c in list(zip(rr, cc))[:12]:
    print("pix", r, c, "got", got[r, c], "want", want[r, c], "seam", dbg["seam_dist"][r, c], "edge", dbg["edge_dist"][r, c])
    for f, d in enumerate(dbg["faces"]):
        sel = (d["rows"] == r) & (d["cols"] == c)
        if sel.any():
            i = np.nonzero(sel)[0][0]
            print("   face", f, "ix,iy", d["ix"][i], d["iy"][i], "flow", flows[f][d["iy"][i], d["ix"][i]])
