"""Stub of `skimage.transform.resize` for running the reference offline (TEST INFRASTRUCTURE ONLY).

The reference only ever calls resize(img, (H, W), preserve_range=True) with the image's own size
(src/projection_icosahedron.py:374-377: the `tuple != list` comparison is always True, so it runs on
every face).  A same-size, preserve_range resize is the identity on values and converts to float64.
"""
import numpy as np


def resize(image, output_shape, preserve_range=False, **kwargs):
    assert tuple(image.shape[:2]) == tuple(int(v) for v in output_shape[:2]), \
        "stub resize only supports the identity (same-size) case the reference uses"
    assert preserve_range
    return np.asarray(image).astype(np.float64)
