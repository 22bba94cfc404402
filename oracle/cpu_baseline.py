"""CPU timing of the oracle port for bench.py (TEST INFRASTRUCTURE ONLY -- cpu_baseline and
--impl reference legs).  The reference is pure Python (numpy/scipy), cannot travel to the GPU box, and
has nothing to compile, so the CPU arm is this port of it (`kind: "port"`); oracle/tif_oracle.py is
pinned to the live reference by tests/test_oracle_golden.py.

The reference itself is single-threaded (it loops the 20 faces serially).  `pair_parallel` gives it
every host core the only way the algorithm allows without changing its arithmetic: the 20 faces of
each stage are independent, so they are farmed out to a process pool and blended in face order.
"""
import multiprocessing as mp
import os
import time

import numpy as np

from . import synth, tif_oracle

_shared = {}


def _init_worker(src, tar, flows, erp_h, wt, pad, blend):
    os.environ["OMP_NUM_THREADS"] = "1"
    _shared.update(src=src, tar=tar, flows=flows, erp_h=erp_h, wt=wt, pad=pad, blend=blend,
                   src64=src.astype(np.float64), tar64=tar.astype(np.float64))


def _resample_task(args):
    which, face = args
    img = _shared["src"] if which == 0 else _shared["tar"]
    return tif_oracle.erp2ico_image(img, _shared["wt"], _shared["pad"], True, faces=[face])[0].astype(np.uint8)


def _stitch_task(face):
    s = _shared
    fr = tif_oracle.stitch_face(face, s["flows"][face], s["erp_h"], s["pad"], s["src64"], s["tar64"], True, s["blend"])
    return {k: fr[k] for k in ("r", "c", "fu", "fv", "w")}


def pair_serial(src, tar, flows, erp_h, wt, pad, blend):
    """resample(src) + resample(tar) + stitch, one core, exactly the reference's call sequence
    (flow_estimate.py:176-185 without DIS)."""
    t0 = time.perf_counter()
    tif_oracle.erp2ico_image(src, wt, pad, True)
    tif_oracle.erp2ico_image(tar, wt, pad, True)
    tif_oracle.ico2erp_flow(flows, erp_h, pad, src, tar, True, blend)
    return time.perf_counter() - t0


class ParallelPair:
    """One frame pair per step, faces spread over `cores` worker processes."""

    def __init__(self, erp_h, wt, pad, blend, cores=None, seed=0):
        self.cores = cores or os.cpu_count() or 1
        self.erp_h, self.wt, self.pad, self.blend = erp_h, wt, pad, blend
        self.src, self.tar = synth.synth_images(erp_h, seed)
        self.flows = synth.synth_flows(wt)
        ctx = mp.get_context("fork")
        self.pool = ctx.Pool(self.cores, initializer=_init_worker,
                             initargs=(self.src, self.tar, self.flows, erp_h, wt, pad, blend))

    def step(self):
        t0 = time.perf_counter()
        self.pool.map(_resample_task, [(w, f) for w in (0, 1) for f in range(20)], chunksize=1)
        # polar faces cover ~2.2x the pixels of the others: submit them first
        order = list(range(0, 5)) + list(range(15, 20)) + list(range(5, 15))
        parts = dict(zip(order, self.pool.map(_stitch_task, order, chunksize=1)))
        tif_oracle.blend_faces([parts[f] for f in range(20)], self.erp_h, True)
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()
