"""B200-native tangent-image hot path of 360OpticalFlow-TangentImages.

The directory name is not a Python identifier; load it with `__graft_entry__.load_package()` (alias
`tangentflow_b200`).  `install_dropin()` registers the modules under the reference's bare module names
(`projection_icosahedron`, `flow_warp`, `flow_postproc`, `gnomonic_projection`, `spherical_coordinates`,
`polygon`) so that reference code such as flow_estimate.PanoOpticalFlow picks them up unchanged.
"""
import sys

from . import _lib  # noqa: F401
from . import polygon, spherical_coordinates, gnomonic_projection  # noqa: F401

__all__ = ["projection_icosahedron", "flow_warp", "flow_postproc", "gnomonic_projection",
           "spherical_coordinates", "polygon", "install_dropin", "sharding"]

_LAZY = ("projection_icosahedron", "projection_cubemap", "projection", "flow_warp", "flow_postproc", "flow_evaluate", "flow_io", "flow_vis",
         "_geometry", "_ops", "sharding", "pipeline")


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod = importlib.import_module("." + name, __name__)
        globals()[name] = mod
        return mod
    raise AttributeError(name)


def install_dropin():
    """Make `import projection_icosahedron` (etc.) resolve to this package, as the reference's scripts expect."""
    for name in ("projection_icosahedron", "projection_cubemap", "projection", "flow_evaluate", "flow_io", "flow_vis", "flow_warp", "flow_postproc", "gnomonic_projection",
                 "spherical_coordinates", "polygon"):
        sys.modules[name] = __getattr__(name) if name in _LAZY else globals()[name]
