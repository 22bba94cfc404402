"""Demo pair + OpenCV DIS flows: GPU vs oracle normwarp stitch, worst pixels with their per-face terms."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2
import __graft_entry__ as entry
from oracle import tif_oracle

entry.build()
pkg = entry.load_package()
g = np.load(os.path.join(ROOT, "tests", "golden", "summary_demo_pair_h640.npz"))
src, tar = g["image_src"], g["image_tar"]
erp_h, WT, PAD = 640, 480, 0.3
pico = pkg.projection_icosahedron
fs, ft = pico.erp2ico_image(src, WT, PAD, True), pico.erp2ico_image(tar, WT, PAD, True)
dis = cv2.DISOpticalFlow_create(cv2.DISOPTICAL_FLOW_PRESET_MEDIUM)
flows = [dis.calc(cv2.cvtColor(a[:, :, :3].astype(np.uint8), cv2.COLOR_BGR2GRAY), cv2.cvtColor(b[:, :, :3].astype(np.uint8), cv2.COLOR_BGR2GRAY), None).astype(np.float32)
         for a, b in zip(fs, ft)]
print("flow abs max", max(float(np.abs(f).max()) for f in flows))
img_s, img_t = src.astype(np.float64), tar.astype(np.float64)
res = [tif_oracle.stitch_face(f, flows[f], erp_h, PAD, img_s, img_t, True, "normwarp") for f in range(20)]
want, dbg = tif_oracle.blend_faces(res, erp_h, True, return_debug=True)
got = pico.ico2erp_flow(flows, erp_h, PAD, src, tar, True, "normwarp")
excuse = (dbg["seam_dist"] < 1e-5) | (dbg["edge_dist"] < 1e-3)
err = np.abs(got - want).max(axis=2)
print("max", err.max(), "bad", int(((err > 1e-3) & ~excuse).sum()), "median", np.median(err), "p99.99", np.quantile(err, 0.9999))
for fidx, fr in enumerate(res):
    d = fr["w"]
print("face gmeans (oracle): n/a; weights min", min(float(fr["w"].min()) for fr in res))
order = np.argsort(-(err * ~excuse).ravel())[:10]
for o in order:
    r, c = divmod(int(o), 2 * erp_h)
    print("r %d c %d err %.3e got %s want %s spread %.3f edge %.3e" % (r, c, err[r, c], got[r, c], want[r, c], dbg["fu_spread"][r, c], dbg["edge_dist"][r, c]))
    for f, fr in enumerate(res):
        sel = (fr["r"] == r) & (fr["c"] == c)
        if sel.any():
            i = np.nonzero(sel)[0][0]
            print("     face %2d fu %.4f fv %.4f w %.6e tx %.4f ty %.4f flow %s" % (f, fr["fu"][i], fr["fv"][i], fr["w"][i], fr["tx"][i], fr["ty"][i], flows[f][fr["iy"][i], fr["ix"][i]]))
