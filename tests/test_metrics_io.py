"""Rows f3 / f4 of SURVEY.md section 8: flow metrics and the .flo format -- oracle vs the live reference (CPU, build
container), CUDA vs the oracle (GPU), file format round trip."""
import importlib
import os

import numpy as np
import pytest

from oracle import ref_shim, tif_oracle


def flows(h, seed=0):
    rng = np.random.default_rng(seed)
    gt = rng.normal(0, 6, (h, 2 * h, 2))
    gt[::5, ::7] = 0.0                      # exact zeros: unavailable pixels
    gt[3, 4, 0] = np.nan
    gt[6, 9] = [3e7, 1.0]                   # too large
    gt[7, 9] = [3e7, 3e7]                   # the precedence quirk keeps this one
    gt[:, -3:, 0] += 2 * h                  # end points across the seam
    ev = gt + rng.normal(0, 0.7, gt.shape)
    mask = rng.uniform(size=(h, 2 * h)) > 0.2
    return gt, ev, mask


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not mounted (GPU box)")
def test_oracle_metrics_against_live_reference():
    ref_shim.load_reference()
    fe = importlib.import_module("flow_evaluate")
    gt, ev, mask = flows(48)
    for spherical in (False, True):
        for m in (None, mask):
            with np.errstate(invalid="ignore"):
                want = (fe.EPE(gt.copy(), ev.copy(), spherical, m), fe.RMSE(gt.copy(), ev.copy(), spherical, m),
                        fe.AAE(gt.copy(), ev.copy(), spherical, m))
                got = tif_oracle.flow_metrics(gt, ev, spherical, m)
            assert np.allclose(got, want, rtol=1e-13, atol=0), (spherical, got, want)
            e_ref, ok_ref = fe.EPE_mat(gt.copy(), ev.copy(), spherical, m)
            e_or, ok_or = tif_oracle.epe_mat(gt, ev, spherical, m)
            assert np.array_equal(ok_ref, ok_or) and np.array_equal(e_ref, e_or)
    assert np.array_equal(fe.available_pixel(gt.copy(), mask), tif_oracle.available_pixel(gt, mask))
    # the six-metric summary: every metric zeroes unusable pixels of both flows in place, the later ones see that
    for m in (None, mask):
        ga, ea, gb, eb = gt.copy(), ev.copy(), gt.copy(), ev.copy()
        want = fe.opticalflow_metric(ga, ea, None, m, verbose=False)
        got = tif_oracle.opticalflow_metric(gb, eb, m)
        assert np.allclose(got, want, rtol=1e-13, atol=0) and np.array_equal(ga, gb) and np.array_equal(ea, eb)


def test_flo_round_trip(pkg, tmp_path):
    fio = pkg.flow_io
    flow = np.random.default_rng(1).normal(size=(13, 26, 2))
    path = str(tmp_path / "x.flo")
    fio.flow_write(flow, path)
    raw = open(path, "rb").read()
    assert raw[:4] == b"PIEH" and np.frombuffer(raw[4:12], np.int32).tolist() == [26, 13] and len(raw) == 12 + 13 * 26 * 8
    back = fio.read_flow_flo(path)
    assert back.dtype == np.float32 and np.array_equal(back, flow.astype(np.float32))
    with pytest.raises(ValueError):
        fio.write_flow_flo(flow, str(tmp_path / "x.bin"))
    if ref_shim.reference_available():
        ref_shim.load_reference()
        rio = importlib.import_module("flow_io")
        rio.write_flow_flo(flow, str(tmp_path / "r.flo"))
        assert open(str(tmp_path / "r.flo"), "rb").read() == raw
        assert np.array_equal(rio.read_flow_flo(path), back)


@pytest.mark.gpu
@pytest.mark.parametrize("h", [48, 512])
def test_gpu_metrics(pkg, h):
    fe = pkg.flow_evaluate
    gt, ev, mask = flows(h)
    for spherical in (False, True):
        for m in (None, mask):
            want = tif_oracle.flow_metrics(gt, ev, spherical, m)
            g1, e1 = gt.copy(), ev.copy()
            got = (fe.EPE(g1, e1, spherical, m), fe.RMSE(gt.copy(), ev.copy(), spherical, m), fe.AAE(gt.copy(), ev.copy(), spherical, m))
            assert np.allclose(got, want, rtol=1e-11, atol=0), (spherical, got, want)
            e_got, ok_got = fe.EPE_mat(gt.copy(), ev.copy(), spherical, m)
            e_want, ok_want = tif_oracle.epe_mat(gt, ev, spherical, m)
            assert np.array_equal(ok_got, ok_want) and np.abs(e_got - e_want).max() < 1e-9
            assert np.array_equal(g1[~ok_want], np.zeros_like(g1[~ok_want]))        # inputs zeroed in place like the reference
    a_got, ok = fe.AAE_mat(gt.copy(), ev.copy(), False, mask)
    a_want, ok_w = tif_oracle.aae_mat(gt, ev, mask)
    assert np.array_equal(ok, ok_w) and np.abs(a_got - a_want).max() < 1e-12
    assert np.array_equal(fe.available_pixel(gt, mask), tif_oracle.available_pixel(gt, mask))
    assert fe.EPE(gt.copy(), ev.copy()) == fe.EPE(gt.copy(), ev.copy())          # deterministic reduction
    # an integer mask enters through `index_valid & of_mask` (src/flow_evaluate.py:48): only its lowest bit counts
    rng = np.random.default_rng(5)
    imask = rng.choice(np.array([0, 1, 2, 254, 255], dtype=np.uint8), size=mask.shape)
    assert np.array_equal(fe.available_pixel(gt, imask), fe.available_pixel(gt, (imask & 1).astype(bool)))
    assert not fe.available_pixel(gt, np.full(mask.shape, 254, dtype=np.uint8)).any()
    with pytest.raises(TypeError):
        fe.available_pixel(gt, mask.astype(np.float64))
    assert np.isnan(fe.EPE(gt.copy(), ev.copy(), False, np.zeros(mask.shape, dtype=bool)))        # empty set: nan, no exception
    with pytest.raises(AttributeError):
        fe.AAE_mat(gt.copy(), ev.copy(), True)
    for m in (None, mask):
        ga, ea, gb, eb = gt.copy(), ev.copy(), gt.copy(), ev.copy()
        got = fe.opticalflow_metric(ga, ea, None, m, verbose=False)
        want = tif_oracle.opticalflow_metric(gb, eb, m)
        assert np.allclose(got, want, rtol=1e-11, atol=0) and np.array_equal(ga, gb) and np.array_equal(ea, eb)


@pytest.mark.gpu
def test_gpu_metric_folder(pkg, tmp_path):
    """The folder driver: .flo files in, one CSV row of six metrics per file."""
    fe, fio = pkg.flow_evaluate, pkg.flow_io
    of_dir, gt_dir = str(tmp_path / "of") + "/", str(tmp_path / "gt") + "/"
    import os
    os.makedirs(of_dir), os.makedirs(gt_dir)
    want = {}
    for i in range(2):
        gt, ev, _ = flows(48, seed=20 + i)
        fio.flow_write(ev, of_dir + "%04d_flow.flo" % i)
        fio.flow_write(gt, gt_dir + "%04d_flow.flo" % i)
        g32, e32 = gt.astype(np.float32), ev.astype(np.float32)
        want["%04d_flow.flo" % i] = tif_oracle.opticalflow_metric(tif_oracle.erp_of_wraparound(g32), tif_oracle.erp_of_wraparound(e32), None)
    fe.opticalflow_metric_folder(of_dir, gt_dir, result_csv_filename="r.csv")
    rows = [r.split(",") for r in open(of_dir + "r.csv").read().strip().splitlines()]
    assert rows[0] == ["flo_filename", "AAE", "EPE", "RMS", "SAAE", "SEPE", "SRMS"] and len(rows) == 3
    for r in rows[1:]:
        assert np.allclose([float(v) for v in r[1:]], want[r[0]], rtol=1e-6)
