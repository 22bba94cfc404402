"""Host face table and geometry-plan management.

The face table (20 faces x a few doubles) and the separable trigonometric tables (H + F*W entries)
are evaluated here in float64 with numpy, in the reference's own operation order
(src/projection_icosahedron.py:19-160, :194-208, :285-305; src/spherical_coordinates.py:92-100), and
handed to `tif_plan_create`.  Everything per pixel happens on the device.  Plans are cached per
(padding, H, Wt, level, device) and stay resident in HBM.
"""
import ctypes as C
import threading

import numpy as np
import torch

from . import _lib
from . import gnomonic_projection as gp
from . import polygon
from . import spherical_coordinates as sc

NUM_FACES_LEVEL0 = 20


def tangent_image_height(tangent_image_width):
    """src/projection_icosahedron.py:194."""
    return int((tangent_image_width / 2.0) / np.tan(np.radians(30.0)) + 0.5)


def _level0_vertices(face):
    """Tangent point and the three spherical vertices of a level-0 face, reference order
    (src/projection_icosahedron.py:43-104; table in SURVEY.md A.1)."""
    r_circ = np.sin(2 * np.pi / 5.0)
    r_in = np.sqrt(3) / 12.0 * (3 + np.sqrt(5))
    r_mid = np.cos(np.pi / 5.0)
    step = 2.0 * np.pi / 5.0
    lat = np.arctan(0.5)
    band, i = divmod(face, 5)
    phi_top = np.pi / 2 - np.arccos(r_in / r_circ)
    if band == 0:
        theta_0 = - np.pi + step / 2.0 + i * step
        phi_0 = phi_top
        verts = ((-np.pi + i * step, lat),
                 (-np.pi + np.pi * 2.0 / 5.0 / 2.0 + i * step, np.pi / 2.0),
                 (-np.pi + (i + 1) * step, lat))
    elif band == 1:
        theta_0 = - np.pi + step / 2.0 + i * step
        phi_0 = np.pi / 2.0 - np.arccos(r_in / r_circ) - 2 * np.arccos(r_in / r_mid)
        verts = ((-np.pi + i * step, lat),
                 (-np.pi + (i + 1) * step, lat),
                 (-np.pi + step / 2.0 + i * step, -lat))
    elif band == 2:
        theta_0 = - np.pi + i * step
        phi_0 = -(np.pi / 2.0 - np.arccos(r_in / r_circ) - 2 * np.arccos(r_in / r_mid))
        verts = ((- np.pi - step / 2.0 + i * step, -lat),
                 (-np.pi + i * step, lat),
                 (- np.pi + step / 2.0 + i * step, -lat))
    elif band == 3:
        theta_0 = - np.pi + i * step
        phi_0 = - phi_top
        verts = ((- np.pi - step / 2.0 + i * step, -lat),
                 (- np.pi + step / 2.0 + i * step, -lat),
                 (- np.pi + i * step, -np.pi / 2.0))
    else:
        raise ValueError("icosahedron face index %r out of range 0..19" % (face,))
    return theta_0, phi_0, verts


def get_icosahedron_parameters(triangle_index, padding_size=0.0):
    """Reference-compatible face dictionary (src/projection_icosahedron.py:19-160)."""
    theta_0, phi_0, verts = _level0_vertices(triangle_index)
    flat = [gp.gnomonic_projection(t, p, theta_0, phi_0) for t, p in verts]
    padded = polygon.enlarge_polygon(flat, padding_size)
    sph = [list(gp.reverse_gnomonic_projection(p[0], p[1], theta_0, phi_0)) for p in padded]
    padded_a = np.array(padded)
    mid_x = np.sort(padded_a[:, 0])[1]
    mid_y = np.sort(padded_a[:, 1])[1]
    probe = np.append(np.array(sph), [gp.reverse_gnomonic_projection(mid_x, mid_y, theta_0, phi_0)], axis=0)
    polar = triangle_index <= 4 or triangle_index >= 15
    if polar and padding_size > 0.0:
        theta_range = [-np.pi, np.pi]
    else:
        theta_range = [np.amin(probe[:, 0]), np.amax(probe[:, 0])]
    if triangle_index <= 4:
        phi_range = [np.pi / 2.0, np.amin(probe[:, 1])]
    elif triangle_index >= 15:
        phi_range = [np.amax(probe[:, 1]), -np.pi / 2.0]
    else:
        phi_range = [np.amax(probe[:, 1]), np.amin(probe[:, 1])]
    return {"tangent_point": [theta_0, phi_0], "triangle_points_tangent": padded,
            "triangle_points_sph": sph, "availied_ERP_area": theta_range + phi_range}


def erp_bounding_box(availied_erp_area, erp_height):
    """Integer ERP bounding box (col_start, col_stop, row_start, row_stop), src/projection_icosahedron.py:299-305."""
    col0, row0 = sc.sph2erp(availied_erp_area[0], availied_erp_area[2], erp_height, sph_modulo=False)
    col1, row1 = sc.sph2erp(availied_erp_area[1], availied_erp_area[3], erp_height, sph_modulo=False)

    def first(v):
        return int(v) if int(v) > 0 else int(v - 0.5)

    def last(v):
        return int(v + 0.5) if int(v) > 0 else int(v)

    return first(col0), last(col1), first(row0), last(row1)


class HostGeometry:
    """Face structs + tables for one (padding, H, Wt) configuration of the 20-face icosahedron."""

    def __init__(self, padding_size, erp_height, tangent_width):
        self.padding_size = float(padding_size)
        self.erp_h = int(erp_height)
        self.erp_w = 2 * self.erp_h
        self.tangent_w = int(tangent_width)
        self.tangent_h = tangent_image_height(self.tangent_w)
        n = NUM_FACES_LEVEL0
        self.n_faces = n
        self.faces = (_lib.TifFace * n)()
        self.params = []
        h, w, wt, ht = self.erp_h, self.erp_w, self.tangent_w, self.tangent_h
        self.gnom_x = np.zeros((n, wt), dtype=np.float64)
        self.gnom_y = np.zeros((n, ht), dtype=np.float64)
        # spherical_coordinates.py:92-100 evaluated on the column / row index vectors
        cols = np.arange(w, dtype=np.float64)
        rows = np.arange(h, dtype=np.float64)
        theta = cols * (2 * np.pi / w) + np.pi / w - np.pi
        phi = -(rows * (np.pi / h) + np.pi / h * 0.5) + 0.5 * np.pi
        theta = np.where(theta == np.pi, -np.pi, theta)
        phi = np.where(phi == -0.5 * np.pi, 0.5 * np.pi, phi)
        self.row_sincos = np.ascontiguousarray(np.stack((np.sin(phi), np.cos(phi)), axis=1))
        self.col_sincos = np.zeros((n, w, 2), dtype=np.float64)
        offset = 0
        for f in range(n):
            par = get_icosahedron_parameters(f, self.padding_size)
            self.params.append(par)
            tri = np.array(par["triangle_points_tangent"])
            xmin, xmax = np.amin(tri[:, 0], axis=0), np.amax(tri[:, 0], axis=0)
            ymin, ymax = np.amin(tri[:, 1], axis=0), np.amax(tri[:, 1], axis=0)
            theta_0, phi_0 = par["tangent_point"]
            F = self.faces[f]
            F.theta0, F.phi0 = theta_0, phi_0
            F.sin_phi0, F.cos_phi0 = np.sin(phi_0), np.cos(phi_0)
            for k in range(3):
                F.tri[k][0], F.tri[k][1] = tri[k, 0], tri[k, 1]
            F.xmin, F.xmax, F.ymin, F.ymax = xmin, xmax, ymin, ymax
            F.g2p_x = (wt - 1.0) / (xmax - xmin + 0.0 * 2.0)           # gnomonic_projection.py:141
            F.g2p_y = (ht - 1.0) / (ymax - ymin + 0.0 * 2.0)           # :145
            F.p2g_x = (wt - 1.0) / (abs(xmax - xmin) + 0.0 * 2.0)      # :188
            F.p2g_y = (ht - 1.0) / (abs(ymax - ymin) + 0.0 * 2.0)      # :192
            F.eps_stitch = abs(xmax - xmin) / (2 * wt)                 # projection_icosahedron.py:296
            F.eps_image = (xmax - xmin) / (wt)                         # :215
            F.col_start, F.col_stop, F.row_start, F.row_stop = erp_bounding_box(par["availied_ERP_area"], h)
            F.tangent_h = ht
            F.pix_offset = offset
            offset += ht * wt
            self.gnom_x[f] = np.linspace(xmin, xmax, num=wt, endpoint=True)     # :207
            self.gnom_y[f] = np.linspace(ymax, ymin, num=ht, endpoint=True)     # :208
            delta = theta - theta_0                                             # gnomonic_projection.py:36
            self.col_sincos[f, :, 0] = np.sin(delta)
            self.col_sincos[f, :, 1] = np.cos(delta)
        self.tangent_pixels = offset

    def tables(self):
        dp = C.POINTER(C.c_double)
        t = _lib.TifPlanTables()
        self._keep = [np.ascontiguousarray(a) for a in
                      (self.gnom_x, self.gnom_y, self.row_sincos, self.col_sincos)]
        t.gnom_x, t.gnom_y, t.row_sincos, t.col_sincos = (a.ctypes.data_as(dp) for a in self._keep)
        return t


class Plan:
    """Device-resident geometry plan (wraps TifPlan*)."""

    def __init__(self, padding_size, erp_height, tangent_width, device):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.TifError("no CUDA device: the tangent-image hot path has no CPU fallback")
        self.device = torch.device(device)
        self.host = HostGeometry(padding_size, erp_height, tangent_width)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            status = self.lib.tif_plan_create(self.host.faces, self.host.n_faces, self.host.erp_h, self.host.erp_w,
                                              self.host.tangent_w, C.byref(self.host.tables()), stream,
                                              C.byref(handle))
        _lib.check(status, "tif_plan_create")
        self.handle = handle
        info = _lib.TifPlanInfo()
        _lib.check(self.lib.tif_plan_info(self.handle, C.byref(info)), "tif_plan_info")
        self.n_faces = info.n_faces
        self.erp_h, self.erp_w = info.erp_h, info.erp_w
        self.tangent_w, self.tangent_h = info.tangent_w, self.host.tangent_h
        self.tangent_pixels = info.tangent_pixels
        self.max_faces_per_pixel = info.max_faces_per_pixel
        self.accepted_unique = info.accepted_unique
        self.referenced_tangent_pixels = info.referenced_tangent_pixels
        self.device_bytes = info.device_bytes

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.tif_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- parity exports --------------------------------------------------------------------
    def export_stage1(self):
        """(erp_x, erp_y) float64 [P] and inside flag uint8 [P] of every tangent pixel."""
        with torch.cuda.device(self.device):
            ex = torch.empty(self.tangent_pixels, dtype=torch.float64, device=self.device)
            ey = torch.empty_like(ex)
            inside = torch.empty(self.tangent_pixels, dtype=torch.uint8, device=self.device)
            _lib.check(self.lib.tif_plan_export_stage1(self.handle, ex.data_ptr(), ey.data_ptr(), inside.data_ptr(),
                                                       torch.cuda.current_stream().cuda_stream), "tif_plan_export_stage1")
        return ex, ey, inside

    def export_stage2(self):
        """count uint8 [H,W] and packed entries int32 [K,H,W] (face<<25 | mult<<22 | iy<<11 | ix)."""
        with torch.cuda.device(self.device):
            count = torch.empty((self.erp_h, self.erp_w), dtype=torch.uint8, device=self.device)
            entries = torch.empty((self.max_faces_per_pixel, self.erp_h, self.erp_w), dtype=torch.int32, device=self.device)
            _lib.check(self.lib.tif_plan_export_stage2(self.handle, count.data_ptr(), entries.data_ptr(),
                                                       torch.cuda.current_stream().cuda_stream), "tif_plan_export_stage2")
        return count, entries


_plans = {}
_plans_lock = threading.Lock()


def get_plan(padding_size, erp_height, tangent_width, device=None):
    """Cached plan for a configuration on a device (default: current CUDA device)."""
    if device is None:
        if not torch.cuda.is_available():
            raise _lib.TifError("no CUDA device: the tangent-image hot path has no CPU fallback")
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = (float(padding_size), int(erp_height), int(tangent_width), device.index)
    with _plans_lock:
        plan = _plans.get(key)
        if plan is None:
            plan = Plan(padding_size, erp_height, tangent_width, device)
            _plans[key] = plan
    return plan


def clear_plans():
    with _plans_lock:
        _plans.clear()
