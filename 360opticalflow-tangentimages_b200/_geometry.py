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


def _unit(v):
    return v / np.linalg.norm(v)


def get_subdivided_parameters(face_index, padding_size=0.0, boundary_samples=256):
    """Face dictionary of the level-1 subdivision (80 faces) -- an EXTENSION, the reference stops at 20 faces
    (SURVEY.md Appendix E).  Built from the reference's own primitives so that nothing new enters the per-pixel
    maths: each level-0 triangle (unpadded spherical vertices of projection_icosahedron.py:43-104) is split at its
    edge midpoints (normalised to the sphere); children (v0,m01,m20), (m01,v1,m12), (m20,m12,v2), (m01,m12,m20);
    face id = 4 * parent + child; tangent point = normalised centroid; padding = padding_size / 2.
    """
    parent, child = divmod(int(face_index), 4)
    _, _, verts = _level0_vertices(parent)
    v = [sc.sph2car(np.float64(t), np.float64(p)) for t, p in verts]          # +x right, +y down, +z forward
    m01, m12, m20 = _unit(v[0] + v[1]), _unit(v[1] + v[2]), _unit(v[2] + v[0])
    tri3d = [(v[0], m01, m20), (m01, v[1], m12), (m20, m12, v[2]), (m01, m12, m20)][child]
    centre = sc.car2sph(np.array([_unit(tri3d[0] + tri3d[1] + tri3d[2])]))[0]
    theta_0, phi_0 = centre[0], centre[1]
    corners = sc.car2sph(np.array(tri3d))
    flat = [gp.gnomonic_projection(corners[k, 0], corners[k, 1], theta_0, phi_0) for k in range(3)]
    flat = [[p[0], p[1]] for p in flat]
    if not polygon.is_clockwise(flat):
        flat = flat[::-1]
    padded = polygon.enlarge_polygon(flat, padding_size / 2.0)
    sph = [list(gp.reverse_gnomonic_projection(p[0], p[1], theta_0, phi_0)) for p in padded]
    # ERP bounding box from a densely sampled boundary of the padded triangle
    bx, by = [], []
    for k in range(3):
        a, b = padded[k], padded[(k + 1) % 3]
        t = np.linspace(0.0, 1.0, boundary_samples, endpoint=False)
        bx.append(a[0] + (b[0] - a[0]) * t)
        by.append(a[1] + (b[1] - a[1]) * t)
    b_theta, b_phi = gp.reverse_gnomonic_projection(np.concatenate(bx), np.concatenate(by), theta_0, phi_0)
    area = [np.amin(b_theta), np.amax(b_theta), np.amax(b_phi), np.amin(b_phi)]
    tiny = 1e-9
    for pole_phi in (np.pi / 2.0, -np.pi / 2.0):
        if np.sin(phi_0) * np.sin(pole_phi) <= 0.0:           # pole on the far hemisphere of this tangent plane
            continue
        px, py, _ = gp.gnomonic_projection(theta_0, pole_phi, theta_0, phi_0)
        if polygon.inside_polygon_2d(np.array([[px, py]]), padded, on_line=True, eps=tiny)[0]:
            if pole_phi > 0:
                area = [-np.pi, np.pi, np.pi / 2.0, np.amin(b_phi)]
            else:
                area = [-np.pi, np.pi, np.amax(b_phi), -np.pi / 2.0]
    return {"tangent_point": [theta_0, phi_0], "triangle_points_tangent": padded,
            "triangle_points_sph": sph, "availied_ERP_area": area}


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
    """Face structs + tables for one (padding, H, Wt, level) configuration: the reference's 20-face icosahedron
    (level 0) or its 80-face subdivision (level 1, extension; per-face tangent heights, ragged stack)."""

    def __init__(self, padding_size, erp_height, tangent_width, level=0):
        if level not in (0, 1):
            raise ValueError("subdivision level must be 0 or 1")
        self.level = int(level)
        self.padding_size = float(padding_size)
        self.erp_h = int(erp_height)
        self.erp_w = 2 * self.erp_h
        self.tangent_w = int(tangent_width)
        self.tangent_h = tangent_image_height(self.tangent_w)          # level 0; level 1 heights are per face
        n = NUM_FACES_LEVEL0 * (4 ** self.level)
        self.n_faces = n
        self.faces = (_lib.TifFace * n)()
        self.params = []
        h, w, wt, ht = self.erp_h, self.erp_w, self.tangent_w, self.tangent_h
        self.gnom_x = np.zeros((n, wt), dtype=np.float64)
        gnom_y_rows = []
        self.face_heights = []
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
            par = (get_icosahedron_parameters(f, self.padding_size) if self.level == 0
                   else get_subdivided_parameters(f, self.padding_size))
            self.params.append(par)
            tri = np.array(par["triangle_points_tangent"])
            xmin, xmax = np.amin(tri[:, 0], axis=0), np.amax(tri[:, 0], axis=0)
            ymin, ymax = np.amin(tri[:, 1], axis=0), np.amax(tri[:, 1], axis=0)
            theta_0, phi_0 = par["tangent_point"]
            if self.level == 1:      # bounding rectangle of the padded triangle at the tangent width
                ht = int(wt * (ymax - ymin) / (xmax - xmin) + 0.5)
            self.face_heights.append(ht)
            F = self.faces[f]
            F.theta0, F.phi0 = theta_0, phi_0
            F.sin_phi0, F.cos_phi0 = np.sin(phi_0), np.cos(phi_0)
            for k in range(3):
                F.tri[k][0], F.tri[k][1] = tri[k, 0], tri[k, 1]
            F.n_vertices, F.n_inner = 3, 0
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
            gnom_y_rows.append(np.linspace(ymax, ymin, num=ht, endpoint=True))  # :208
            delta = theta - theta_0                                             # gnomonic_projection.py:36
            self.col_sincos[f, :, 0] = np.sin(delta)
            self.col_sincos[f, :, 1] = np.cos(delta)
        self.tangent_pixels = offset
        self.gnom_y = np.concatenate(gnom_y_rows)

    def tables(self):
        dp = C.POINTER(C.c_double)
        t = _lib.TifPlanTables()
        self._keep = [np.ascontiguousarray(a) for a in
                      (self.gnom_x, self.gnom_y, self.row_sincos, self.col_sincos)]
        t.gnom_x, t.gnom_y, t.row_sincos, t.col_sincos = (a.ctypes.data_as(dp) for a in self._keep)
        return t


CUBEMAP_TANGENT_POINTS = ((np.pi / 2.0, 0.0), (-np.pi / 2.0, 0.0), (0.0, -np.pi / 2.0), (0.0, np.pi / 2.0),
                          (0.0, 0.0), (-np.pi, 0.0))          # +x, -x, +y, -y, +z, -z (projection_cubemap.py:33-40)


def cubemap_face_erp_range(padding_size=0.0):
    """Per face [[theta, phi] x TL, TR, BR, BL] of the padded square in the ERP image
    (src/projection_cubemap.py:86-118)."""
    pbc = 1.0 + padding_size
    out = np.zeros((6, 4, 2), dtype=float)
    for index, (theta_0, phi_0) in enumerate(CUBEMAP_TANGENT_POINTS):
        if index in (0, 1, 4, 5):
            theta_min, _ = gp.reverse_gnomonic_projection(-pbc, 0, theta_0, phi_0)
            theta_max, _ = gp.reverse_gnomonic_projection(pbc, 0, theta_0, phi_0)
            _, phi_max = gp.reverse_gnomonic_projection(0, pbc, theta_0, phi_0)
            _, phi_min = gp.reverse_gnomonic_projection(0, -pbc, theta_0, phi_0)
        else:
            _, phi_ = gp.reverse_gnomonic_projection(-pbc, pbc, theta_0, phi_0)
            theta_min, theta_max = -np.pi, np.pi
            phi_max, phi_min = (phi_, -np.pi / 2) if index == 2 else (np.pi / 2, phi_)
        out[index] = [[theta_min, phi_max], [theta_max, phi_max], [theta_max, phi_min], [theta_min, phi_min]]
    return out


def cubemap_erp_bounding_box(face_range, erp_height):
    """Integer ERP bounding box of a cubemap face, src/projection_cubemap.py:202-236 (face_meshgrid), with its
    special cases at the image borders (including the plain `if` on the last row test)."""
    face_range = np.asarray(face_range)
    theta_min, theta_max = face_range[:, 0].min(), face_range[:, 0].max()
    phi_min, phi_max = face_range[:, 1].min(), face_range[:, 1].max()
    x_min, y_max = sc.sph2erp(theta_min, phi_min, erp_height, False)
    x_max, y_min = sc.sph2erp(theta_max, phi_max, erp_height, False)
    x_min = 0 if theta_min == -np.pi else (int(x_min) if int(x_min) > 0 else int(x_min - 0.5))
    x_max = erp_height * 2 - 1 if theta_max == np.pi else (int(x_max + 0.5) if int(x_max) > 0 else int(x_max))
    y_min = 0 if phi_max == np.pi * 0.5 else (int(y_min) if int(y_min) >= 0 else int(y_min - 0.5))
    if phi_min == -np.pi * 0.5:
        y_max = erp_height - 1
    y_max = int(y_max + 0.5) if int(y_max) > 0 else int(y_max)
    return x_min, x_max, y_min, y_max


class CubemapGeometry:
    """Face structs + tables of the 6-face cubemap stage (src/projection_cubemap.py) for one (padding, H, S)."""

    def __init__(self, padding_size, erp_height, face_size):
        self.level, self.kind = 0, _lib.PLAN_CUBEMAP
        self.padding_size = float(padding_size)
        self.erp_h, self.erp_w = int(erp_height), 2 * int(erp_height)
        self.tangent_w = self.tangent_h = int(face_size)
        n = self.n_faces = 6
        h, w, s = self.erp_h, self.erp_w, self.tangent_w
        pbc = 1.0 + self.padding_size
        self.faces = (_lib.TifFace * n)()
        self.face_heights = [s] * n
        cols = np.arange(w, dtype=np.float64)
        rows = np.arange(h, dtype=np.float64)
        theta = cols * (2 * np.pi / w) + np.pi / w - np.pi
        phi = -(rows * (np.pi / h) + np.pi / h * 0.5) + 0.5 * np.pi
        theta = np.where(theta == np.pi, -np.pi, theta)
        phi = np.where(phi == -0.5 * np.pi, 0.5 * np.pi, phi)
        self.row_sincos = np.ascontiguousarray(np.stack((np.sin(phi), np.cos(phi)), axis=1))
        self.col_sincos = np.zeros((n, w, 2), dtype=np.float64)
        self.gnom_x = np.tile(np.linspace(-pbc, pbc, s), (n, 1))                 # projection_cubemap.py:160
        self.gnom_y = np.tile(np.linspace(pbc, -pbc, s), n)                      # :161
        ranges = cubemap_face_erp_range(self.padding_size)
        square = [[-pbc, pbc], [pbc, pbc], [pbc, -pbc], [-pbc, -pbc]]             # :294-295
        unit = [[-1, 1], [1, 1], [1, -1], [-1, -1]]                                # projection.py:90-91 (pbc = 1)
        for f, (theta_0, phi_0) in enumerate(CUBEMAP_TANGENT_POINTS):
            F = self.faces[f]
            F.theta0, F.phi0 = theta_0, phi_0
            F.sin_phi0, F.cos_phi0 = np.sin(phi_0), np.cos(phi_0)
            for k in range(4):
                F.tri[k][0], F.tri[k][1] = square[k]
                F.inner[k][0], F.inner[k][1] = unit[k]
            F.n_vertices, F.n_inner = 4, 4
            F.xmin, F.xmax, F.ymin, F.ymax = -pbc, pbc, -pbc, pbc
            F.g2p_x = (s - 1.0) / (pbc - (-pbc) + 0.0 * 2.0)
            F.g2p_y = (s - 1.0) / (pbc - (-pbc) + 0.0 * 2.0)
            F.p2g_x = (s - 1.0) / (abs(pbc - (-pbc)) + 0.0 * 2.0)
            F.p2g_y = (s - 1.0) / (abs(pbc - (-pbc)) + 0.0 * 2.0)
            F.eps_stitch = 1e-4                     # default eps of inside_polygon_2d (polygon.py:59)
            F.eps_image = 1e-4
            F.eps_inner = 1e-7                      # projection.py:92
            x0, x1, y0, y1 = cubemap_erp_bounding_box(ranges[f], h)
            F.col_start, F.col_stop, F.row_start, F.row_stop = x0, x1, y0, y1
            F.tangent_h = s
            F.pix_offset = f * s * s
            delta = theta - theta_0
            self.col_sincos[f, :, 0] = np.sin(delta)
            self.col_sincos[f, :, 1] = np.cos(delta)
        self.tangent_pixels = n * s * s

    tables = HostGeometry.tables


class Plan:
    """Device-resident geometry plan (wraps TifPlan*)."""

    def __init__(self, padding_size, erp_height, tangent_width, device, level=0, kind=_lib.PLAN_ICOSAHEDRON):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.TifError("no CUDA device: the tangent-image hot path has no CPU fallback")
        self.device = torch.device(device)
        self.kind = int(kind)
        if kind == _lib.PLAN_CUBEMAP:
            self.host = CubemapGeometry(padding_size, erp_height, tangent_width)
        else:
            self.host = HostGeometry(padding_size, erp_height, tangent_width, level)
        self.level = int(level)
        self.face_heights = list(self.host.face_heights)
        self.face_offsets = [int(self.host.faces[f].pix_offset) for f in range(self.host.n_faces)]
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            status = self.lib.tif_plan_create(self.host.faces, self.host.n_faces, self.host.erp_h, self.host.erp_w,
                                              self.host.tangent_w, C.byref(self.host.tables()), self.kind, stream,
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
        self.stage1_patches = info.stage1_patches
        self.stage1_staged_patches = info.stage1_staged_patches
        self.stage1_tma_patches = info.stage1_tma_patches
        self.stage1_tma_box = (info.stage1_tma_box_rows, tuple(info.stage1_tma_box_bytes))

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


def get_plan(padding_size, erp_height, tangent_width, device=None, level=0, kind=_lib.PLAN_ICOSAHEDRON):
    """Cached plan for a configuration on a device (default: current CUDA device)."""
    if device is None:
        if not torch.cuda.is_available():
            raise _lib.TifError("no CUDA device: the tangent-image hot path has no CPU fallback")
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = (float(padding_size), int(erp_height), int(tangent_width), device.index, int(level), int(kind))
    with _plans_lock:
        plan = _plans.get(key)
        if plan is None:
            plan = Plan(padding_size, erp_height, tangent_width, device, level, kind)
            _plans[key] = plan
    return plan


def clear_plans():
    """Drop every cached plan; their device tables are freed as soon as no caller holds a plan any more."""
    import sys
    with _plans_lock:
        _plans.clear()
    ops = sys.modules.get(__package__ + "._ops")
    if ops is not None:
        ops.unregister_all()
