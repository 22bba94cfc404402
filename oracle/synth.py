"""Synthetic inputs.  TEST INFRASTRUCTURE ONLY (bench.py has its own device generator)."""
import numpy as np


def tangent_height(tangent_w):
    return int((tangent_w / 2.0) / np.tan(np.radians(30.0)) + 0.5)


def synth_images(erp_h, seed=0, smooth=False):
    """src, tar uint8 [H, 2H, 3]; `src` is drawn first.  smooth=True gives a band-limited pair."""
    erp_w = 2 * erp_h
    rng = np.random.default_rng(seed)
    if not smooth:
        src = rng.integers(0, 256, (erp_h, erp_w, 3), dtype=np.uint8)
        tar = rng.integers(0, 256, (erp_h, erp_w, 3), dtype=np.uint8)
        return src, tar
    yy, xx = np.meshgrid(np.arange(erp_h), np.arange(erp_w), indexing="ij")
    out = []
    for k in range(2):
        chans = []
        for ch in range(3):
            a, b, p = rng.uniform(1, 6), rng.uniform(1, 4), rng.uniform(0, 6.28)
            chans.append(127.5 + 100.0 * np.sin(2 * np.pi * a * xx / erp_w + p) * np.cos(np.pi * b * yy / erp_h + 0.3 * k))
        out.append(np.clip(np.stack(chans, -1), 0, 255).astype(np.uint8))
    return out[0], out[1]


def synth_flows(tangent_w=480, faces=20, u_offset=0.0, tangent_h=None):
    """Analytic tangent flows, float32 [Ht, Wt, 2] per face."""
    if tangent_h is None:
        tangent_h = tangent_height(tangent_w)
    yy, xx = np.meshgrid(np.arange(tangent_h), np.arange(tangent_w), indexing="ij")
    flows = []
    for f in range(faces):
        u = 4.0 * np.sin(2 * np.pi * (xx / tangent_w + f / float(faces))) + 1.5 + u_offset
        v = 3.0 * np.cos(2 * np.pi * (yy / tangent_h - f / float(faces))) - 0.5
        flows.append(np.stack((u, v), -1).astype(np.float32))
    return flows


def synth(erp_h, seed=0, tangent_w=480, smooth=False, u_offset=0.0):
    src, tar = synth_images(erp_h, seed, smooth)
    return src, tar, synth_flows(tangent_w, 20, u_offset)
