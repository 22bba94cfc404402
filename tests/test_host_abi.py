"""Host-side logic and the C-ABI surface (no GPU needed): face table, library exports, struct layout,
loud failure without a device, batch sharding over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_face_table_matches_reference(pkg, golden):
    geo = pkg._geometry
    g = golden("face_table")
    for pi_, pad in enumerate(g["pads"]):
        for f in range(20):
            par = geo.get_icosahedron_parameters(f, float(pad))
            assert np.array_equal(np.array(par["tangent_point"]), g["tangent_point"][pi_, f])
            assert np.array_equal(np.array(par["triangle_points_tangent"]), g["tri"][pi_, f])
            assert np.array_equal(np.array(par["availied_ERP_area"]), g["area"][pi_, f])
            for hi, h in enumerate(g["bbox_heights"]):
                assert tuple(g["bbox"][pi_, hi, f]) == geo.erp_bounding_box(par["availied_ERP_area"], int(h))


def test_host_geometry_tables(pkg):
    hg = pkg._geometry.HostGeometry(0.3, 64, 32)
    assert hg.tangent_h == 28 and hg.tangent_pixels == 20 * 28 * 32
    assert hg.row_sincos.shape == (64, 2) and hg.col_sincos.shape == (20, 128, 2)
    assert [hg.faces[f].pix_offset for f in range(3)] == [0, 28 * 32, 2 * 28 * 32]
    assert hg.faces[0].col_start == -1 and hg.faces[0].col_stop == 128 and hg.faces[0].row_start == -1
    assert pkg._geometry.tangent_image_height(480) == 416


def test_library_exports_every_declared_symbol(pkg):
    header = open(os.path.join(ROOT, "include", "tif_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(tif_[a-z0-9_]+)\s*\(", body))
    assert len(declared) >= 16
    lib = ctypes.CDLL(pkg._lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "libtif_b200.so does not export %s" % name
    assert declared == set(pkg._lib.SYMBOLS), "binding and header disagree: %s" % (declared ^ set(pkg._lib.SYMBOLS))


def test_struct_layout_and_status_strings(pkg):
    lib = pkg._lib.load()
    assert lib.tif_abi_version() == pkg._lib.ABI_VERSION == 5
    sizes = [ctypes.c_int32() for _ in range(3)]
    assert lib.tif_struct_sizes(*[ctypes.byref(s) for s in sizes]) == 0
    assert [s.value for s in sizes] == [ctypes.sizeof(pkg._lib.TifFace), ctypes.sizeof(pkg._lib.TifPlanTables),
                                        ctypes.sizeof(pkg._lib.TifPlanInfo)]
    assert lib.tif_status_string(0) == b"TIF_OK" and lib.tif_status_string(-4) == b"TIF_ERR_RANGE"
    # argument validation happens before any CUDA call
    assert lib.tif_plan_create(None, 0, 0, 0, 0, None, 0, None, None) == -1
    assert b"invalid argument" in lib.tif_last_error()


def test_host_primitives_match_oracle(pkg):
    from oracle import tif_oracle
    rng = np.random.default_rng(5)
    th, ph = rng.uniform(-3.1, 3.1, 500), rng.uniform(-1.5, 1.5, 500)
    x, y, _ = pkg.gnomonic_projection.gnomonic_projection(th, ph, 0.3, -0.9)
    ox, oy = tif_oracle.gnomonic_forward(th, ph, 0.3, -0.9)
    assert np.array_equal(x, ox) and np.array_equal(y, oy)
    t2, p2 = pkg.gnomonic_projection.reverse_gnomonic_projection(x, y, 0.3, -0.9)
    o2, q2 = tif_oracle.gnomonic_reverse(ox, oy, 0.3, -0.9)
    assert np.array_equal(t2, o2) and np.array_equal(p2, q2)
    tri = [[-1.0, -0.6], [0.0, 1.2], [1.0, -0.6]]
    pts = rng.uniform(-1.5, 1.5, (2000, 2))
    assert np.array_equal(pkg.polygon.inside_polygon_2d(pts, tri, True, 1e-2),
                          tif_oracle.inside_polygon(pts[:, 0], pts[:, 1], tri, 1e-2))
    ix, iy = pkg.gnomonic_projection.gnomonic2pixel(pts[:, 0], pts[:, 1], 0.0, 48, 40, [-1.5, 1.5, -1.5, 1.5])
    assert ix.min() >= 0 and ix.max() <= 47 and iy.min() >= 0 and iy.max() <= 39
    ex, ey = pkg.spherical_coordinates.sph2erp(th, ph, 64, True)
    ox, oy = tif_oracle.sph2erp(th, ph, 64, True)
    assert np.array_equal(ex, ox) and np.array_equal(ey, oy)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_fails_loudly_without_a_gpu(pkg):
    img = np.zeros((32, 64, 3), np.uint8)
    with pytest.raises(pkg._lib.TifError):
        pkg.projection_icosahedron.erp2ico_image(img, 16, 0.3)
    with pytest.raises(pkg._lib.TifError):
        pkg.flow_warp.warp_backward(img, np.zeros((32, 64, 2)))
    with pytest.raises(pkg._lib.TifError):
        pkg.flow_postproc.erp_of_wraparound(np.zeros((32, 64, 2)))
    with pytest.raises(Exception):
        torch.ops.tangentflow.flow_warp_meshgrid(torch.zeros(1, 4, 8), torch.zeros(1, 4, 8))


def test_shard_range(pkg):
    sh = pkg.sharding
    for n in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [sh.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sh.shard_sizes(n, world)
    with pytest.raises(ValueError):
        sh.shard_range(4, 2, 2)


@pytest.mark.parametrize("n_items", [6, 5])
def test_gather_over_gloo_world_size_2(n_items):
    """The N>1 path of bench.py / sharding.py on CPU: two gloo ranks shard a batch and gather it."""
    script = os.path.join(ROOT, "tests", "_gloo_gather_worker.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(29500 + n_items), script, str(n_items)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "GATHER_OK" in res.stdout


def test_level1_face_table_matches_oracle(pkg):
    """Config #5 (80 faces, extension): the product's host table and the oracle's independent restatement agree
    bitwise, the per-face heights are SURVEY.md Appendix E's, and level 0 of the same class is the reference table."""
    from oracle import tif_oracle
    geo = pkg._geometry
    for f in range(80):
        a = geo.get_subdivided_parameters(f, 0.3)
        b = tif_oracle.subdivided_face(f, 0.3)
        assert a["tangent_point"][0] == b["theta_0"] and a["tangent_point"][1] == b["phi_0"]
        assert np.array_equal(np.array(a["triangle_points_tangent"]), b["tri"])
        assert np.array_equal(np.array(a["availied_ERP_area"]), np.array(b["area"]))
    hg = geo.HostGeometry(0.3, 2048, 480, level=1)
    assert hg.n_faces == 80 and sorted(set(hg.face_heights)) == [346, 416, 485, 502] and sum(hg.face_heights) == 34980
    assert hg.gnom_y.shape == (34980,) and hg.tangent_pixels == 34980 * 480
    hg0 = geo.HostGeometry(0.3, 64, 32, level=0)
    assert hg0.n_faces == 20 and set(hg0.face_heights) == {28}


@pytest.mark.skipif(not os.path.isfile("/root/reference/src/flow_estimate.py"), reason="/root/reference not mounted (GPU box)")
def test_install_dropin_binds_the_reference_caller():
    """After install_dropin() the reference's own pipeline module (src/flow_estimate.py:1-8, imported unmodified
    from /root/reference) resolves its bare-name imports to this package: PanoOpticalFlow.estimate would run the
    B200 path.  Run in a subprocess so that the bare module names do not leak into the test process."""
    code = r"""
import sys
sys.path.insert(0, %r)
import __graft_entry__ as entry
entry.build()
pkg = entry.load_package()
pkg.install_dropin()
sys.path.insert(0, %r)                       # colorama / skimage / matplotlib stubs (not installed on this image)
sys.path.append("/root/reference/src")       # appended: the product modules win for every name they provide
import flow_estimate
assert flow_estimate.proj_ico is pkg.projection_icosahedron, flow_estimate.proj_ico
assert flow_estimate.proj_cm is pkg.projection_cubemap
assert flow_estimate.flow_warp is pkg.flow_warp and flow_estimate.projection is pkg.projection
assert flow_estimate.flow_postproc is pkg.flow_postproc
assert flow_estimate.spherical_coordinates is pkg.spherical_coordinates
est = flow_estimate.PanoOpticalFlow()
assert est.face_blending_method_ico == "normwarp"
for name in ("erp2ico_image", "ico2erp_flow", "get_icosahedron_parameters"):
    assert getattr(flow_estimate.proj_ico, name).__module__.startswith("tangentflow_b200"), name
print("dropin ok")
""" % (ROOT, os.path.join(ROOT, "oracle", "_stubs"))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "dropin ok" in res.stdout, res.stdout + res.stderr
