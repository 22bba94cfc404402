"""Import the UNMODIFIED reference (yuanmingze/360OpticalFlow-TangentImages) from /root/reference.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this file.  It exists so that
(a) `oracle/make_golden.py` can generate the committed fixtures in tests/golden/ from the live
reference, and (b) tests that run in the build container (where /root/reference is mounted) can pin
`oracle/tif_oracle.py` against the real thing.  /root/reference does not exist on the GPU box, so
every caller must check `reference_available()` first.

The reference cannot be imported as-is on this image because
  * numpy >= 1.24 removed the aliases np.float / np.int / np.bool that it uses
    (gnomonic_projection.py:41,81,143,147; polygon.py:148,156; projection.py:41,48-50; ...);
  * colorama (logger.py:5), skimage (projection_icosahedron.py:5) and matplotlib (flow_vis.py:1)
    are not installed (stubs in oracle/_stubs/).
The reference modules import each other by bare name with src/ on sys.path (main.py:3-13), so they
are loaded by bare name here too and handed back in a dict; product modules of the same bare names
are never placed on sys.path.
"""
import importlib
import os
import sys

import numpy as np

REFERENCE_SRC = os.environ.get("TIF_REFERENCE_SRC", "/root/reference/src")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_stubs")

_REF_MODULES = (
    "logger", "polygon", "spherical_coordinates", "gnomonic_projection", "flow_postproc",
    "pointcloud_utils", "flow_warp", "projection", "projection_icosahedron",
)

_cache = {}


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_SRC, "projection_icosahedron.py"))


def load_reference():
    """Return {bare module name: module} of the live reference."""
    if _cache:
        return _cache
    if not reference_available():
        raise RuntimeError("reference sources not found at %s" % REFERENCE_SRC)
    for name, typ in (("float", float), ("int", int), ("bool", bool)):
        if not hasattr(np, name):
            setattr(np, name, typ)
    for p in (_STUBS, REFERENCE_SRC):
        if p not in sys.path:
            sys.path.insert(0, p)
    for name in _REF_MODULES:
        mod = importlib.import_module(name)
        origin = os.path.abspath(getattr(mod, "__file__", ""))
        if not origin.startswith(os.path.abspath(REFERENCE_SRC)):
            raise RuntimeError("module %s resolved to %s, not the reference" % (name, origin))
        _cache[name] = mod
    import logging
    logging.disable(logging.WARNING)   # the reference warns on every gnomonic2pixel call
    return _cache
