"""CPU oracle: numpy restatement of the reference's tangent-image hot path.

TEST INFRASTRUCTURE ONLY -- the checker, never the thing measured or shipped.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The product package (360opticalflow-tangentimages_b200/) never does.

Parity pin: the reference holds no golden vectors or tests of its own, so the
pin is the reference itself run live in the build container: `oracle/make_golden.py` imports the
unmodified reference through `oracle/ref_shim.py` and writes tests/golden/*.npz; tests/test_oracle_*.py
check this restatement against those fixtures (everywhere) and against the live reference (where
/root/reference exists).  Third-party arithmetic on the path -- scipy.ndimage.map_coordinates(order=1)
(scipy 1.18.1 here, unpinned in the reference's requirements.txt:5) -- is restated in
`map_coordinates_order1` from scipy's ni_interpolation.c and pinned against the installed scipy.

Every function cites the reference file:line (relative to /root/reference/src) it follows.
All arithmetic is float64 with one IEEE rounding per operation, in the reference's operation order,
so face membership and pixel indices are bit-identical to the reference on the same host.
"""
import numpy as np

PI = np.pi
NUM_FACES = 20

# ------------------------------------------------------------------------------------------------
# scipy.ndimage.map_coordinates(order=1) -- third-party; call sites projection_icosahedron.py:241,
# 333,336; projection.py:52-53; flow_warp.py:84,86.  Restated from scipy/ndimage/src/ni_interpolation.c
# (NI_GeometricTransform + map_coordinate + get_spline_interpolation_weights).
# ------------------------------------------------------------------------------------------------


def _wrap_coordinate(c, n):
    """scipy mode='wrap' (NI_EXTEND_WRAP): periodic with period n-1, NOT n."""
    c = np.array(c, dtype=np.float64, copy=True)
    if n <= 1:
        c[...] = 0.0
        return c
    sz = float(n - 1)
    neg = c < 0
    if neg.any():
        cn = c[neg]
        c[neg] = cn + sz * (np.trunc(-cn / sz) + 1.0)
    big = c > (n - 1)
    if big.any():
        cb = c[big]
        c[big] = cb - sz * np.trunc(cb / sz)
    return c


def map_coordinates_order1(image, ys, xs, mode, cval=0.0):
    """Bilinear sample of a 2-D array at float coordinates with scipy's boundary rules.

    mode: 'wrap' (period n-1), 'constant' (any coordinate <0 or >n-1 -> cval outright),
    'nearest' (clamp).  Output dtype = input dtype; for uint8 the double accumulator t is
    converted as `t>0 ? t+0.5 : 0`, clamped to [0,255], truncated (round-half-up).
    Accumulation order = scipy's: taps (y0,x0),(y0,x1),(y1,x0),(y1,x1), each p*wy*wx, summed in double.
    """
    image = np.asarray(image)
    H, W = image.shape
    ys = np.asarray(ys, dtype=np.float64)
    xs = np.asarray(xs, dtype=np.float64)
    shape = ys.shape
    ys = ys.ravel()
    xs = xs.ravel()
    outside = np.zeros(ys.shape, dtype=bool)
    if mode == "wrap":
        ys = _wrap_coordinate(ys, H)
        xs = _wrap_coordinate(xs, W)
    elif mode == "constant":
        outside = (ys < 0) | (ys > H - 1) | (xs < 0) | (xs > W - 1)
        ys = np.where(outside, 0.0, ys)
        xs = np.where(outside, 0.0, xs)
    elif mode == "nearest":
        ys = np.clip(ys, 0.0, H - 1.0)
        xs = np.clip(xs, 0.0, W - 1.0)
    else:
        raise ValueError(mode)
    y0f = np.floor(ys)
    x0f = np.floor(xs)
    y0 = y0f.astype(np.int64)
    x0 = x0f.astype(np.int64)
    fy = ys - y0f
    fx = xs - x0f
    wy0 = 1.0 - fy
    wy1 = 1.0 - wy0     # scipy: last weight = 1 - sum(others)
    wx0 = 1.0 - fx
    wx1 = 1.0 - wx0
    # the +1 tap falls outside only when the coordinate is exactly n-1 (weight 0): scipy mirrors it
    y1 = np.where(y0 + 1 >= H, max(H - 2, 0), y0 + 1)
    x1 = np.where(x0 + 1 >= W, max(W - 2, 0), x0 + 1)
    img = image.astype(np.float64, copy=False)
    t = img[y0, x0] * wy0 * wx0
    t = t + img[y0, x1] * wy0 * wx1
    t = t + img[y1, x0] * wy1 * wx0
    t = t + img[y1, x1] * wy1 * wx1
    if mode == "constant":
        t = np.where(outside, float(cval), t)
    if image.dtype == np.uint8:
        t = np.where(t > 0, t + 0.5, 0.0)
        t = np.clip(t, 0.0, 255.0)
        return t.astype(np.uint8).reshape(shape)
    if image.dtype == np.float32:
        return t.astype(np.float32).reshape(shape)
    return t.reshape(shape)


# ------------------------------------------------------------------------------------------------
# geometry primitives
# ------------------------------------------------------------------------------------------------

def gnomonic_forward(theta, phi, theta_0, phi_0):
    """gnomonic_projection.py:18-56 (x, y only; the hemisphere mask is computed and discarded by
    the stitcher, projection_icosahedron.py:317)."""
    cos_c = np.sin(phi_0) * np.sin(phi) + np.cos(phi_0) * np.cos(phi) * np.cos(theta - theta_0)
    zeros = cos_c == 0
    if np.any(zeros):
        cos_c = np.where(zeros, np.finfo(float).eps, cos_c)
    x = np.cos(phi) * np.sin(theta - theta_0) / cos_c
    y = (np.cos(phi_0) * np.sin(phi) - np.sin(phi_0) * np.cos(phi) * np.cos(theta - theta_0)) / cos_c
    if np.any(zeros):
        x = np.where(zeros, 0.0, x)
        y = np.where(zeros, 0.0, y)
    return x, y


def gnomonic_reverse(x, y, theta_0, phi_0):
    """gnomonic_projection.py:59-95.  theta is continuous (may leave [-pi,pi)); rho==0 -> (theta_0, phi_0)."""
    rho = np.sqrt(x ** 2 + y ** 2)
    zeros = rho == 0
    if np.any(zeros):
        rho = np.where(zeros, np.finfo(float).eps, rho)
    c = np.arctan2(rho, 1)
    phi_ = np.arcsin(np.cos(c) * np.sin(phi_0) + (y * np.sin(c) * np.cos(phi_0)) / rho)
    theta_ = theta_0 + np.arctan2(x * np.sin(c), rho * np.cos(phi_0) * np.cos(c) - y * np.sin(phi_0) * np.sin(c))
    if np.any(zeros):
        phi_ = np.where(zeros, phi_0, phi_)
        theta_ = np.where(zeros, theta_0, theta_)
    return theta_, phi_


def erp_sph_modulo(theta, phi):
    """spherical_coordinates.py:55-61."""
    return np.remainder(theta + PI, 2 * PI) - PI, -(np.remainder(-phi + 0.5 * PI, PI) - 0.5 * PI)


def sph_coord_modulo(theta, phi):
    """spherical_coordinates.py:234-239 (note +pi/2 -> -pi/2)."""
    return np.remainder(theta + PI, PI * 2.0) - PI, np.remainder(phi + 0.5 * PI, PI) - 0.5 * PI


def sph2erp(theta, phi, erp_h, sph_modulo=False):
    """spherical_coordinates.py:104-125."""
    if sph_modulo:
        theta, phi = erp_sph_modulo(theta, phi)
    erp_w = 2 * erp_h
    return (theta + PI) / (2.0 * PI / erp_w) - 0.5, (-phi + 0.5 * PI) / (PI / erp_h) - 0.5


def erp2sph_axes(cols, rows, erp_h):
    """spherical_coordinates.py:64-101 evaluated on separable column / row index vectors
    (every operation there is elementwise, so a 1-D evaluation is bitwise the 2-D one)."""
    width = erp_h * 2
    theta = cols * (2 * PI / width) + PI / width - PI
    phi = -(rows * (PI / erp_h) + PI / erp_h * 0.5) + 0.5 * PI
    theta = np.where(theta == PI, -PI, theta)
    phi = np.where(phi == -0.5 * PI, 0.5 * PI, phi)
    return theta, phi


def inside_polygon(px, py, poly, eps):
    """polygon.py:59-118 with on_line=True: ray cast with inclusive y range and eps-wide edges.
    px, py: arrays of equal shape; poly: [n,2].  Returns inside | online."""
    inside = np.zeros(px.shape, dtype=bool)
    online = np.zeros(px.shape, dtype=bool)
    n = len(poly)
    for k in range(n):
        x1, y1 = poly[k][0], poly[k][1]
        x2, y2 = poly[(k + 1) % n][0], poly[(k + 1) % n][1]
        t = (py >= min(y1, y2)) & (py <= max(y1, y2)) & (px <= max(x1, x2))
        if not t.any():
            continue
        if abs(y1 - y2) <= eps:
            t = t & (px >= min(x1, x2))
            ix = px
        else:
            ix = (py - y1) * (x2 - x1) / (y2 - y1) + x1
        online |= t & (np.abs(px - ix) <= eps)
        inside ^= t & (px <= ix)
    return inside | online


def _find_intersection(p1, p2, p3, p4):
    """polygon.py:8-40."""
    dx12 = p2[0] - p1[0]
    dy12 = p2[1] - p1[1]
    dx34 = p4[0] - p3[0]
    dy34 = p4[1] - p3[1]
    den = (dy12 * dx34 - dx12 * dy34)
    if den == 0:
        return None
    t1 = ((p1[0] - p3[0]) * dy34 + (p3[1] - p1[1]) * dx34) / den
    return [p1[0] + dx12 * t1, p1[1] + dy12 * t1]


def enlarge_polygon(pts, offset):
    """polygon.py:121-167: push every edge outward by `offset` (absolute gnomonic units) and intersect."""
    out = []
    n = len(pts)
    for j in range(n):
        i = (j - 1) % n
        k = (j + 1) % n
        v1 = np.array([pts[j][0] - pts[i][0], pts[j][1] - pts[i][1]], float)
        v1 = v1 / np.linalg.norm(v1) * offset
        n1 = [-v1[1], v1[0]]
        pij1 = [pts[i][0] + n1[0], pts[i][1] + n1[1]]
        pij2 = [pts[j][0] + n1[0], pts[j][1] + n1[1]]
        v2 = np.array([pts[k][0] - pts[j][0], pts[k][1] - pts[j][1]], float)
        v2 = v2 / np.linalg.norm(v2) * offset
        n2 = [-v2[1], v2[0]]
        pjk1 = [pts[j][0] + n2[0], pts[j][1] + n2[1]]
        pjk2 = [pts[k][0] + n2[0], pts[k][1] + n2[1]]
        out.append(_find_intersection(pij1, pij2, pjk1, pjk2))
    return out


# ------------------------------------------------------------------------------------------------
# face table -- projection_icosahedron.py:19-160
# ------------------------------------------------------------------------------------------------

def icosahedron_face(face, padding=0.0, return_vertices=False):
    """Returns dict(theta_0, phi_0, tri [3,2] padded gnomonic triangle, area [th_min, th_max, phi_rowstart, phi_rowstop])."""
    r_circ = np.sin(2 * np.pi / 5.0)
    r_in = np.sqrt(3) / 12.0 * (3 + np.sqrt(5))
    r_mid = np.cos(np.pi / 5.0)
    step = 2.0 * np.pi / 5.0
    at = np.arctan(0.5)
    if 0 <= face <= 4:
        i = face
        theta_0 = - np.pi + step / 2.0 + i * step
        phi_0 = np.pi / 2 - np.arccos(r_in / r_circ)
        v = [(-np.pi + i * step, at),
             (-np.pi + np.pi * 2.0 / 5.0 / 2.0 + i * step, np.pi / 2.0),
             (-np.pi + (i + 1) * step, at)]
    elif 5 <= face <= 9:
        i = face - 5
        theta_0 = - np.pi + step / 2.0 + i * step
        phi_0 = np.pi / 2.0 - np.arccos(r_in / r_circ) - 2 * np.arccos(r_in / r_mid)
        v = [(-np.pi + i * step, at),
             (-np.pi + (i + 1) * step, at),
             (-np.pi + step / 2.0 + i * step, -at)]
    elif 10 <= face <= 14:
        i = face - 10
        theta_0 = - np.pi + i * step
        phi_0 = -(np.pi / 2.0 - np.arccos(r_in / r_circ) - 2 * np.arccos(r_in / r_mid))
        v = [(- np.pi - step / 2.0 + i * step, -at),
             (-np.pi + i * step, at),
             (- np.pi + step / 2.0 + i * step, -at)]
    elif 15 <= face <= 19:
        i = face - 15
        theta_0 = - np.pi + i * step
        phi_0 = - (np.pi / 2 - np.arccos(r_in / r_circ))
        v = [(- np.pi - step / 2.0 + i * step, -at),
             (- np.pi + step / 2.0 + i * step, -at),
             (- np.pi + i * step, -np.pi / 2.0)]
    else:
        raise ValueError(face)
    if return_vertices:
        return v
    tri0 = [gnomonic_forward(th, ph, theta_0, phi_0) for th, ph in v]
    tri = enlarge_polygon(tri0, padding)
    sph = [list(gnomonic_reverse(p[0], p[1], theta_0, phi_0)) for p in tri]
    tri_a = np.array(tri)
    mx = np.sort(tri_a[:, 0])[1]
    my = np.sort(tri_a[:, 1])[1]
    sph_a = np.append(np.array(sph), [gnomonic_reverse(mx, my, theta_0, phi_0)], axis=0)
    if face <= 4 or face >= 15:
        if padding > 0.0:
            th = [-np.pi, np.pi]
        else:
            th = [np.amin(sph_a[:, 0]), np.amax(sph_a[:, 0])]
        if face <= 4:
            area = th + [np.pi / 2.0, np.amin(sph_a[:, 1])]
        else:
            area = th + [np.amax(sph_a[:, 1]), -np.pi / 2.0]
    else:
        area = [np.amin(sph_a[:, 0]), np.amax(sph_a[:, 0]), np.amax(sph_a[:, 1]), np.amin(sph_a[:, 1])]
    return {"theta_0": theta_0, "phi_0": phi_0, "tri": tri_a, "area": area}


def _sph2car(theta, phi):
    """spherical_coordinates.py:151-169: +x right, +y down, +z forward."""
    return np.array([np.cos(phi) * np.sin(theta), -np.sin(phi), np.cos(phi) * np.cos(theta)])


def _car2sph(points):
    """spherical_coordinates.py:128-148: [N,3] cartesian -> [N,2] (theta, phi)."""
    points = np.asarray(points, dtype=np.float64)
    radius = np.linalg.norm(points, axis=1)
    return np.stack((np.arctan2(points[:, 0], points[:, 2]), -np.arcsin(np.divide(points[:, 1], radius))), axis=1)


def _is_clockwise(pts):
    """polygon.py:43-56."""
    total = 0
    for i in range(len(pts)):
        cur, nxt = pts[i], pts[(i + 1) % len(pts)]
        total += (nxt[0] - cur[0]) * (nxt[1] + cur[1])
    return total > 0


def subdivided_face(face, padding=0.0, boundary_samples=256):
    """Level-1 (80-face) subdivision, SURVEY.md Appendix E -- an EXTENSION: the reference has no such table, so
    this is a composition of its primitives and parity for it is pinned only by self-consistency
    (level 0 of the same machinery == the reference).  Same dictionary as icosahedron_face."""
    parent, child = divmod(int(face), 4)
    verts = icosahedron_face(parent, 0.0, return_vertices=True)
    v = [_sph2car(np.float64(t), np.float64(p)) for t, p in verts]
    mid = lambda a, b: (a + b) / np.linalg.norm(a + b)      # noqa: E731
    m01, m12, m20 = mid(v[0], v[1]), mid(v[1], v[2]), mid(v[2], v[0])
    tri3d = [(v[0], m01, m20), (m01, v[1], m12), (m20, m12, v[2]), (m01, m12, m20)][child]
    c = tri3d[0] + tri3d[1] + tri3d[2]
    theta_0, phi_0 = _car2sph(np.array([c / np.linalg.norm(c)]))[0]
    corners = _car2sph(np.array(tri3d))
    flat = []
    for k in range(3):
        x, y = gnomonic_forward(corners[k, 0], corners[k, 1], theta_0, phi_0)
        flat.append([x, y])
    if not _is_clockwise(flat):
        flat = flat[::-1]
    tri = enlarge_polygon(flat, padding / 2.0)
    tri_a = np.array(tri)
    bx, by = [], []
    for k in range(3):
        a, b = tri[k], tri[(k + 1) % 3]
        t = np.linspace(0.0, 1.0, boundary_samples, endpoint=False)
        bx.append(a[0] + (b[0] - a[0]) * t)
        by.append(a[1] + (b[1] - a[1]) * t)
    b_th, b_ph = gnomonic_reverse(np.concatenate(bx), np.concatenate(by), theta_0, phi_0)
    area = [np.amin(b_th), np.amax(b_th), np.amax(b_ph), np.amin(b_ph)]
    for pole in (np.pi / 2.0, -np.pi / 2.0):
        if np.sin(phi_0) * np.sin(pole) <= 0.0:
            continue
        px, py = gnomonic_forward(theta_0, pole, theta_0, phi_0)
        if inside_polygon(np.array([px]), np.array([py]), tri_a, 1e-9)[0]:
            area = [-np.pi, np.pi, np.pi / 2.0, np.amin(b_ph)] if pole > 0 else [-np.pi, np.pi, np.amax(b_ph), -np.pi / 2.0]
    return {"theta_0": theta_0, "phi_0": phi_0, "tri": tri_a, "area": area}


def face_params(face, padding, level=0):
    return icosahedron_face(face, padding) if level == 0 else subdivided_face(face, padding)


def n_faces(level=0):
    return NUM_FACES * 4 ** level


def face_tangent_height(face_par, tangent_w, level=0):
    """projection_icosahedron.py:194 at level 0; bounding rectangle of the padded triangle at level 1."""
    if level == 0:
        return tangent_height(tangent_w)
    tri = face_par["tri"]
    return int(tangent_w * (np.amax(tri[:, 1]) - np.amin(tri[:, 1])) / (np.amax(tri[:, 0]) - np.amin(tri[:, 0])) + 0.5)


def _pyint_start(v):
    return int(v) if int(v) > 0 else int(v - 0.5)


def _pyint_stop(v):
    return int(v + 0.5) if int(v) > 0 else int(v)


def face_erp_bbox(face_par, erp_h):
    """projection_icosahedron.py:299-305: integer (col_start, col_stop, row_start, row_stop); may leave the image."""
    a = face_par["area"]
    cs, rs = sph2erp(a[0], a[2], erp_h, False)
    ce, re = sph2erp(a[1], a[3], erp_h, False)
    return _pyint_start(cs), _pyint_stop(ce), _pyint_start(rs), _pyint_stop(re)


def tangent_height(tangent_w):
    """projection_icosahedron.py:194."""
    return int((tangent_w / 2.0) / np.tan(np.radians(30.0)) + 0.5)


# ------------------------------------------------------------------------------------------------
# stage 1: ERP -> tangent images -- projection_icosahedron.py:163-248
# ------------------------------------------------------------------------------------------------

def erp2ico_sample_coords(face_par, erp_h, tangent_w, tangent_h):
    """Float ERP sample coordinate of every tangent pixel of one face (:203-225)."""
    tri = face_par["tri"]
    xmin, xmax = np.amin(tri[:, 0]), np.amax(tri[:, 0])
    ymin, ymax = np.amin(tri[:, 1]), np.amax(tri[:, 1])
    gx = np.linspace(xmin, xmax, num=tangent_w, endpoint=True)
    gy = np.linspace(ymax, ymin, num=tangent_h, endpoint=True)
    gxv, gyv = np.meshgrid(gx, gy)
    th, ph = gnomonic_reverse(gxv, gyv, face_par["theta_0"], face_par["phi_0"])
    ex, ey = sph2erp(th, ph, erp_h, True)
    return ex, ey, gxv, gyv, (xmin, xmax, ymin, ymax)


def erp2ico_image(erp_image, tangent_w, padding=0.0, full_face_image=True, faces=None, level=0):
    """Returns list of int64 [Ht,Wt,4] (RGB rounded uint8 values, alpha 255/0) like the reference."""
    erp_image = np.asarray(erp_image)
    if erp_image.ndim == 3 and erp_image.shape[2] == 4:
        erp_image = erp_image[:, :, 0:3]
    elif erp_image.ndim == 2:
        erp_image = erp_image[:, :, None]
    erp_h = erp_image.shape[0]
    nch = erp_image.shape[2]
    out = []
    for f in (range(n_faces(level)) if faces is None else faces):
        par = face_params(f, padding, level)
        tangent_h = face_tangent_height(par, tangent_w, level)
        ex, ey, gxv, gyv, (xmin, xmax, ymin, ymax) = erp2ico_sample_coords(par, erp_h, tangent_w, tangent_h)
        inside = np.ones(gxv.shape, dtype=bool)
        if not full_face_image:
            eps = (xmax - xmin) / tangent_w
            inside = inside_polygon(gxv, gyv, par["tri"], eps)
        # gnomonic2pixel (gnomonic_projection.py:141-147) with padding 0 and the padded range
        tx = ((gxv[inside] - xmin + 0.0) * ((tangent_w - 1.0) / (xmax - xmin + 0.0)) + 0.5).astype(np.int64)
        ty = (-(gyv[inside] - ymax - 0.0) * ((tangent_h - 1.0) / (ymax - ymin + 0.0)) + 0.5).astype(np.int64)
        img = np.full([tangent_h, tangent_w, 4 if nch == 3 else nch], 255, dtype=np.int64)
        for ch in range(nch):
            img[ty, tx, ch] = map_coordinates_order1(erp_image[:, :, ch], ey[inside], ex[inside], "wrap", 255)
        if img.shape[2] == 4:
            img[:, :, 3] = 0
            img[ty, tx, 3] = 255
        out.append(img)
    return out


# ------------------------------------------------------------------------------------------------
# stage 2+3: tangent flows -> ERP flow -- projection_icosahedron.py:251-398, "pull" form per face
# ------------------------------------------------------------------------------------------------

def _axis_candidates(start, stop, n):
    """Unique wrapped indices of the integer range [start, stop] with their multiplicities
    (flow_postproc.py:9-14 maps j -> rem(j+0.5, n)-0.5 == j mod n for integer j)."""
    idx = np.remainder(np.arange(start, stop + 1, dtype=np.float64) + 0.5, n) - 0.5
    vals, counts = np.unique(idx.astype(np.int64), return_counts=True)
    return vals, counts


def face_membership(face_par, erp_h, tangent_w, tangent_h):
    """Steps A.3-1..5 of SURVEY.md for one face: candidate rows/cols, accepted mask, nearest tangent pixel.
    Returns rows [R], cols [C], mult [R,C], accepted [R,C] bool, ix [R,C] int64, iy [R,C] int64, range tuple."""
    erp_w = 2 * erp_h
    tri = face_par["tri"]
    th0, ph0 = face_par["theta_0"], face_par["phi_0"]
    xmin, xmax = np.amin(tri[:, 0]), np.amax(tri[:, 0])
    ymin, ymax = np.amin(tri[:, 1]), np.amax(tri[:, 1])
    eps = abs(xmax - xmin) / (2 * tangent_w)
    cs, ce, rs, re = face_erp_bbox(face_par, erp_h)
    cols, mc = _axis_candidates(cs, ce, erp_w)
    rows, mr = _axis_candidates(rs, re, erp_h)
    theta, phi = erp2sph_axes(cols.astype(np.float64), rows.astype(np.float64), erp_h)
    x, y = gnomonic_forward(theta[None, :], phi[:, None], th0, ph0)
    acc = inside_polygon(x, y, tri, eps)
    ix = ((x - xmin + 0.0) * ((tangent_w - 1.0) / (xmax - xmin + 0.0 * 2.0)) + 0.5)
    iy = (-(y - ymax - 0.0) * ((tangent_h - 1.0) / (ymax - ymin + 0.0 * 2.0)) + 0.5)
    with np.errstate(invalid="ignore"):
        ix = np.where(acc, ix, 0.0).astype(np.int64)
        iy = np.where(acc, iy, 0.0).astype(np.int64)
    mult = mr[:, None] * mc[None, :]
    return rows, cols, mult, acc, ix, iy, (xmin, xmax, ymin, ymax), theta


def stitch_face(face, flow, erp_h, padding, img_s, img_t, wrap_around, blending, level=0):
    """One iteration of the reference's face loop (projection_icosahedron.py:282-381) in pull form.
    flow: float32 [Ht,Wt,2]; img_s/img_t: float64 [H,W,3] (normwarp only).
    Returns dict(r, c, fu, fv, w, ix, iy, mult, candidates, accepted, bbox, th_t, tx, ty)."""
    flow = np.asarray(flow)
    tangent_h, tangent_w = flow.shape[:2]
    erp_w = int(erp_h * 2.0)
    par = face_params(face, padding, level)
    th0, ph0 = par["theta_0"], par["phi_0"]
    rows, cols, mult, acc, ix, iy, (xmin, xmax, ymin, ymax), theta_cols = \
        face_membership(par, erp_h, tangent_w, tangent_h)
    rr, cc = np.nonzero(acc)
    r = rows[rr]
    c = cols[cc]
    ixa = ix[rr, cc]
    iya = iy[rr, cc]
    # map_coordinates(order=1, mode='nearest') at integer coordinates == clamped gather (:333-337)
    gi = np.clip(iya, 0, tangent_h - 1)
    gj = np.clip(ixa, 0, tangent_w - 1)
    px = ixa + flow[gi, gj, 0]          # int64 + float32 -> float64
    py = iya + flow[gi, gj, 1]
    # pixel2gnomonic (gnomonic_projection.py:186-194)
    gx = px / ((tangent_w - 1.0) / (abs(xmax - xmin) + 0.0 * 2.0)) + xmin - 0.0
    gy = - py / ((tangent_h - 1.0) / (abs(ymax - ymin) + 0.0 * 2.0)) + ymax + 0.0
    th_t, ph_t = gnomonic_reverse(gx, gy, th0, ph0)
    th_t, ph_t = sph_coord_modulo(th_t, ph_t)
    tx, ty = sph2erp(th_t, ph_t, erp_h, True)
    if wrap_around:
        th_s = np.remainder(theta_cols[cc] + PI, PI * 2.0) - PI
        long_line = np.abs(th_t - th_s) > PI
        minus = long_line & (th_s < 0) & (th_t > 0)
        plus = long_line & (th_s > 0) & (th_t < 0)
        tx = np.where(minus, tx - erp_w, tx)
        tx = np.where(plus, tx + erp_w, tx)
    fu = tx - c
    fv = ty - r
    m = mult[rr, cc]
    if blending == "straightforward":
        w = np.ones(fu.shape, dtype=np.float64)
    elif blending == "normwarp":
        # projection.py:15-61 -- weight = exp(-mean_ch|tar(tx,ty)-src(c,r)| / face-global mean),
        # duplicates of wrapped candidate pixels counted with their multiplicity in the global mean
        d = np.zeros((fu.shape[0], 3), dtype=np.float64)
        for ch in range(3):
            ts = map_coordinates_order1(img_t[:, :, ch], ty, tx, "constant", 255)
            d[:, ch] = np.abs(ts - img_s[r, c, ch])
        mf = m.astype(np.float64)
        gmean = np.sum(d * mf[:, None]) / (3.0 * np.sum(mf))
        w = np.exp(-(np.mean(d, axis=1) / gmean))
    else:
        raise ValueError(blending)
    return {"r": r, "c": c, "fu": fu, "fv": fv, "w": w, "ix": ixa, "iy": iya, "mult": m,
            "candidates": int(mult.sum()), "accepted": int(m.sum()), "bbox": face_erp_bbox(par, erp_h),
            "th_t": th_t, "tx": tx, "ty": ty}


def blend_faces(face_results, erp_h, wrap_around, return_debug=False):
    """Accumulate per-face results in face order, normalise, wrap (projection_icosahedron.py:384-396)."""
    erp_w = int(erp_h * 2.0)
    acc_flow = np.zeros((erp_h, erp_w, 2), dtype=np.float64)
    acc_w = np.zeros((erp_h, erp_w), dtype=np.float64)
    # distance of the closest contributing end point to a discontinuity of the reference's own map:
    # the +-pi seam (u jumps by W without the seam fix) and the hard edge of mode='constant' sampling
    seam_dist = np.full((erp_h, erp_w), np.inf)
    edge_dist = np.full((erp_h, erp_w), np.inf)
    fu_min = np.full((erp_h, erp_w), np.inf)
    fu_max = np.full((erp_h, erp_w), -np.inf)
    debug = []
    for fr in face_results:
        r, c, w = fr["r"], fr["c"], fr["w"]
        acc_flow[r, c, 0] += fr["fu"] * w
        acc_flow[r, c, 1] += fr["fv"] * w
        acc_w[r, c] += w
        if return_debug:
            tx, ty = fr["tx"], fr["ty"]
            np.minimum.at(seam_dist, (r, c), PI - np.abs(fr["th_t"]))
            e = np.minimum(np.minimum(np.abs(tx), np.abs(tx - (erp_w - 1))), np.minimum(np.abs(ty), np.abs(ty - (erp_h - 1))))
            np.minimum.at(edge_dist, (r, c), e)
            np.minimum.at(fu_min, (r, c), fr["fu"])
            np.maximum.at(fu_max, (r, c), fr["fu"])
            debug.append({"rows": r, "cols": c, "ix": fr["ix"], "iy": fr["iy"], "mult": fr["mult"],
                          "candidates": fr["candidates"], "accepted": fr["accepted"], "bbox": fr["bbox"]})
    nz = acc_w != 0.0
    for ch in range(2):
        plane = acc_flow[:, :, ch]
        plane[nz] = plane[nz] / acc_w[nz]
    if wrap_around:
        acc_flow = erp_of_wraparound(acc_flow)
    if return_debug:
        return acc_flow, {"faces": debug, "seam_dist": seam_dist, "edge_dist": edge_dist, "fu_spread": fu_max - fu_min}
    return acc_flow


def ico2erp_flow(tangent_flows, erp_h=None, padding=0.0, image_src=None, image_tar=None,
                 wrap_around=False, blending="straightforward", return_debug=False, level=0):
    """Returns float64 [H, 2H, 2] (and a debug dict when asked)."""
    tangent_h = tangent_flows[0].shape[0]
    if erp_h is None:
        erp_h = int(tangent_h * 2.0)
    erp_h = int(erp_h)
    img_s = img_t = None
    if blending == "normwarp":
        img_s = np.asarray(image_src).astype(np.float64)     # the same-size `resize` (:374-377) is this cast
        img_t = np.asarray(image_tar).astype(np.float64)
    results = [stitch_face(f, tangent_flows[f], erp_h, padding, img_s, img_t, wrap_around, blending, level)
               for f in range(len(tangent_flows))]
    return blend_faces(results, erp_h, wrap_around, return_debug)


# ------------------------------------------------------------------------------------------------
# flow_postproc / flow_warp
# ------------------------------------------------------------------------------------------------

def erp_pixles_modulo(x, y, image_width, image_height):
    """flow_postproc.py:9-14."""
    return np.remainder(x + 0.5, image_width) - 0.5, np.remainder(y + 0.5, image_height) - 0.5


def erp_of_wraparound(erp_flow, of_u_threshold=None):
    """flow_postproc.py:17-44.  Mutates the u plane of its argument (through a view) like the reference."""
    w = erp_flow.shape[1]
    if of_u_threshold is None:
        of_u_threshold = w / 2.0
    u = erp_flow[:, :, 0]
    split = int(w - 1 - of_u_threshold)
    left = np.zeros(u.shape, dtype=bool)
    left[:, 0:split] = True
    sel = left & (u > of_u_threshold)
    u[sel] = u[sel] - w
    right = np.zeros(u.shape, dtype=bool)
    right[:, split:w - 1] = True
    sel = right & (u < -of_u_threshold)
    u[sel] = u[sel] + w
    return np.stack((u, erp_flow[:, :, 1]), axis=2)


def flow_warp_meshgrid(flow_u, flow_v):
    """flow_warp.py:15-51: end points with a SINGLE wrap (not a modulo)."""
    h, w = flow_u.shape
    xa, ya = np.meshgrid(np.linspace(0, w - 1, w), np.linspace(0, h - 1, h))
    eu = xa + flow_u
    ev = ya + flow_v
    eu = np.where(eu >= w - 0.5, eu - w, eu)
    eu = np.where(eu < -0.5, eu + w, eu)
    ev = np.where(ev >= h - 0.5, ev - h, ev)
    ev = np.where(ev < -0.5, ev + h, ev)
    return np.stack((eu, ev))


def warp_backward(image_target, of_forward):
    """flow_warp.py:54-88: out(y,x) = bilinear(img, (y+v, x+u)); 'wrap' for HxWxC, 'constant' 255 for HxW."""
    image_target = np.asarray(image_target)
    h, w = image_target.shape[:2]
    xa, ya = np.meshgrid(np.linspace(0, w - 1, w), np.linspace(0, h - 1, h))
    xn = xa + of_forward[:, :, 0]
    yn = ya + of_forward[:, :, 1]
    out = np.zeros_like(image_target)
    if image_target.ndim == 3:
        for ch in range(image_target.shape[2]):
            out[:, :, ch] = map_coordinates_order1(image_target[:, :, ch], yn, xn, "wrap")
    else:
        out[:, :] = map_coordinates_order1(image_target, yn, xn, "constant", 255)
    return out


# ------------------------------------------------------------------------------------------------
# cubemap stage -- projection_cubemap.py
# ------------------------------------------------------------------------------------------------

CUBEMAP_CENTRES = ((PI / 2.0, 0.0), (-PI / 2.0, 0.0), (0.0, -PI / 2.0), (0.0, PI / 2.0), (0.0, 0.0), (-PI, 0.0))


def cubemap_face_ranges(padding=0.0):
    """projection_cubemap.py:86-118: per face [theta_min, theta_max, phi_max, phi_min] of the padded square."""
    pbc = 1.0 + padding
    out = []
    for index, (theta_0, phi_0) in enumerate(CUBEMAP_CENTRES):
        if index in (0, 1, 4, 5):
            theta_min, _ = gnomonic_reverse(-pbc, 0, theta_0, phi_0)
            theta_max, _ = gnomonic_reverse(pbc, 0, theta_0, phi_0)
            _, phi_max = gnomonic_reverse(0, pbc, theta_0, phi_0)
            _, phi_min = gnomonic_reverse(0, -pbc, theta_0, phi_0)
        else:
            _, phi_ = gnomonic_reverse(-pbc, pbc, theta_0, phi_0)
            theta_min, theta_max = -PI, PI
            if index == 2:
                phi_max, phi_min = phi_, -PI / 2
            else:
                phi_max, phi_min = PI / 2, phi_
        out.append((theta_min, theta_max, phi_max, phi_min))
    return out


def cubemap_erp_bbox(face_range, erp_h):
    """projection_cubemap.py:190-244 (face_meshgrid): integer (x_min, x_max, y_min, y_max), special-cased at the
    image borders.  Note the reference's `if` (not `elif`) on the last branch, kept here."""
    theta_min, theta_max, phi_max, phi_min = face_range
    x_min, y_max = sph2erp(theta_min, phi_min, erp_h, False)
    x_max, y_min = sph2erp(theta_max, phi_max, erp_h, False)
    if theta_min == -PI:
        x_min = 0
    elif int(x_min) > 0:
        x_min = int(x_min)
    else:
        x_min = int(x_min - 0.5)
    if theta_max == PI:
        x_max = erp_h * 2 - 1
    elif int(x_max) > 0:
        x_max = int(x_max + 0.5)
    else:
        x_max = int(x_max)
    if phi_max == PI * 0.5:
        y_min = 0
    elif int(y_min) >= 0:
        y_min = int(y_min)
    else:
        y_min = int(y_min - 0.5)
    if phi_min == -PI * 0.5:
        y_max = erp_h - 1
    if int(y_max) > 0:
        y_max = int(y_max + 0.5)
    else:
        y_max = int(y_max)
    return x_min, x_max, y_min, y_max


def erp2cubemap_image(erp_image, padding=0.0, face_size=None):
    """projection_cubemap.py:127-187: six float64 [S,S,C] faces (uint8 inputs give uint8-rounded values)."""
    erp_image = np.asarray(erp_image)
    erp_h, erp_w, nch = erp_image.shape
    if face_size is None:
        face_size = int(erp_w / 4.0)
    pbc = 1.0 + padding
    out = []
    for theta_0, phi_0 in CUBEMAP_CENTRES:
        x, y = np.meshgrid(np.linspace(-pbc, pbc, face_size), np.linspace(pbc, -pbc, face_size))
        th, ph = gnomonic_reverse(x, y, theta_0, phi_0)
        ex = ((th + PI) / (2 * PI)) * erp_w
        ey = (-ph + PI / 2.0) / PI * erp_h
        ex = np.where(ex < 0, ex + erp_w, ex)
        ex = np.where(ex >= erp_w, ex - erp_w, ex)
        ey = np.where(ey < 0, ey + erp_h, ey)
        ey = np.where(ey >= erp_h, ey - erp_h, ey)
        face = np.zeros((face_size, face_size, nch), dtype=float)
        for ch in range(nch):
            face[:, :, ch] = map_coordinates_order1(erp_image[:, :, ch], ey, ex, "wrap")
        out.append(face)
    return out


def cubemap2erp_flow(cubemap_flows, erp_h=None, padding=0.0, image_src=None, image_tar=None, return_debug=False):
    """projection_cubemap.py:246-357 (face_flows_wraparound=None; the `wrap_around` argument of the reference is
    never read).  Weights: 0.95 / ||tar(end point) - src||_2 when both images are given (projection.py:94-113),
    else 1 inside the unpadded square [-1,1]^2 and 0 outside (:87-93)."""
    size = cubemap_flows[0].shape[0]
    if erp_h is None:
        erp_h = size * 2.0
    erp_h = int(erp_h)
    erp_w = int(erp_h * 2.0)
    acc_flow = np.zeros((erp_h, erp_w, 2), dtype=np.float64)
    acc_w = np.zeros((erp_h, erp_w), dtype=np.float64)
    pbc = 1.0 + padding
    square = np.array([[-pbc, pbc], [pbc, pbc], [pbc, -pbc], [-pbc, -pbc]])
    unit = np.array([[-1, 1], [1, 1], [1, -1], [-1, -1]])
    ranges = cubemap_face_ranges(padding)
    debug = []
    for f in range(6):
        theta_0, phi_0 = CUBEMAP_CENTRES[f]
        x_min, x_max, y_min, y_max = cubemap_erp_bbox(ranges[f], erp_h)
        cols = np.remainder(np.linspace(x_min, x_max, x_max - x_min + 1), erp_w)
        rows = np.remainder(np.linspace(y_min, y_max, y_max - y_min + 1), erp_h)
        theta, phi = erp2sph_axes(cols, rows, erp_h)
        gx, gy = gnomonic_forward(theta[None, :], phi[:, None], theta_0, phi_0)
        acc = inside_polygon(gx, gy, square, 1e-4)
        rr, cc = np.nonzero(acc)
        r = rows[rr].astype(np.int64)
        c = cols[cc].astype(np.int64)
        gxa, gya = gx[rr, cc], gy[rr, cc]
        ix = ((gxa + pbc + 0.0) * ((size - 1.0) / (pbc + pbc + 0.0)) + 0.5).astype(np.int64)
        iy = (-(gya - pbc - 0.0) * ((size - 1.0) / (pbc + pbc + 0.0)) + 0.5).astype(np.int64)
        flow = np.asarray(cubemap_flows[f])
        inr = (ix >= 0) & (ix <= size - 1) & (iy >= 0) & (iy <= size - 1)      # mode='constant', cval=0.0
        u = np.where(inr, flow[np.clip(iy, 0, size - 1), np.clip(ix, 0, size - 1), 0], 0.0)
        v = np.where(inr, flow[np.clip(iy, 0, size - 1), np.clip(ix, 0, size - 1), 1], 0.0)
        px = ix + u
        py = iy + v
        ratio = (size - 1.0) / (abs(pbc + pbc) + 0.0 * 2.0)
        tgx = px / ratio + (-pbc) - 0.0
        tgy = - py / ratio + pbc + 0.0
        th_t, ph_t = gnomonic_reverse(tgx, tgy, theta_0, phi_0)
        tx, ty = sph2erp(th_t, ph_t, erp_h, True)
        fu = tx - cols[cc]
        fv = ty - rows[rr]
        if image_src is not None and image_tar is not None:
            d = np.zeros((fu.shape[0], 3), dtype=np.float64)
            for ch in range(3):
                ts = map_coordinates_order1(np.asarray(image_tar)[:, :, ch], ty, tx, "constant", 255)
                ss = np.asarray(image_src)[r, c, ch]
                # uint8 images: map_coordinates returns uint8 and the reference stores it into a float array
                d[:, ch] = np.absolute(ts.astype(np.float64) - ss.astype(np.float64))
            norm = np.linalg.norm(d, axis=1)
            w = np.ones(fu.shape[0], dtype=np.float64)
            nzw = norm != 0.0
            w[nzw] = 0.95 / norm[nzw]
        else:
            w = inside_polygon(gxa, gya, unit, 1e-7).astype(np.float64)
        acc_flow[r, c, 0] += fu * w
        acc_flow[r, c, 1] += fv * w
        acc_w[r, c] += w
        if return_debug:
            debug.append({"rows": r, "cols": c, "ix": ix, "iy": iy, "bbox": (x_min, x_max, y_min, y_max),
                          "inner": inside_polygon(gxa, gya, unit, 1e-7), "th_t": th_t, "tx": tx, "ty": ty})
    nz = acc_w != 0
    for ch in range(2):
        plane = acc_flow[:, :, ch]
        plane[nz] = plane[nz] / acc_w[nz]
    if return_debug:
        return acc_flow, debug
    return acc_flow


# ------------------------------------------------------------------------------------------------
# global rotation -- flow_warp.py:91-133,190-229, spherical_coordinates.py:128-218,
# projection.py:120-159, pointcloud_utils.py:10-32
# ------------------------------------------------------------------------------------------------

def erp2sph_grid(x, y, erp_h):
    """spherical_coordinates.py:64-101 on full arrays (sph_modulo=False)."""
    width = erp_h * 2
    theta = x * (2 * PI / width) + PI / width - PI
    phi = -(y * (PI / erp_h) + PI / erp_h * 0.5) + 0.5 * PI
    theta = np.where(theta == PI, -PI, theta)
    phi = np.where(phi == -0.5 * PI, 0.5 * PI, phi)
    return theta, phi


def sph2car_grid(theta, phi):
    """spherical_coordinates.py:151-169 with radius 1.0: [3, ...]."""
    return np.stack((1.0 * np.cos(phi) * np.sin(theta), -1.0 * np.sin(phi), 1.0 * np.cos(phi) * np.cos(theta)), axis=0)


def rotation2erp_motion_vector(array_size, rotation_matrix, wraparound=False):
    """spherical_coordinates.py:185-218: ERP flow of a rigid rotation of the sphere, [H,W,2] float64."""
    h, w = array_size
    vx, vy = np.meshgrid(np.linspace(0, w, w, endpoint=False), np.linspace(0, h, h, endpoint=False))
    theta, phi = erp2sph_grid(vx, vy, h)
    xyz = sph2car_grid(theta, phi)
    rot = np.dot(rotation_matrix, xyz.reshape((3, -1)))
    sph = _car2sph(rot.T).T
    rx, ry = sph2erp(sph[0, :], sph[1, :], h, False)
    mx = rx.reshape((h, w)) - vx
    my = ry.reshape((h, w)) - vy
    if wraparound:
        tar_theta = sph[0, :].reshape((h, w))
        long_line = np.abs(tar_theta - theta) > PI
        minus = long_line & (theta < 0) & (tar_theta > 0)
        plus = long_line & (theta > 0) & (tar_theta < 0)
        mx[minus] = mx[minus] - w
        mx[plus] = mx[plus] + w
    return np.stack((mx, my), -1)


def rotate_erp_array(erp_image, rotation_mat):
    """spherical_coordinates.py:172-182."""
    return warp_backward(erp_image, rotation2erp_motion_vector(erp_image.shape[0:2], rotation_mat.T))


def flow_cross_covariance(erp_flow):
    """The 3x3 matrix H = S^T T of flow_warp.flow2rotation_3d (flow_warp.py:103-127) + pointcloud_utils.py:29:
    unit vectors of the pixel centres (S) and of the flow end points (T) over the middle half of the rows."""
    h, w = erp_flow.shape[:2]
    tar = flow_warp_meshgrid(erp_flow[:, :, 0], erp_flow[:, :, 1])
    xa, ya = np.meshgrid(np.linspace(0, w - 1, w), np.linspace(0, h - 1, h))
    st, sp = erp2sph_grid(xa, ya, h)
    tt, tp = erp2sph_grid(tar[0], tar[1], h)
    r0, r1 = int(h * 0.25), int(h * 0.75)
    s3 = sph2car_grid(st[r0:r1], sp[r0:r1])
    t3 = sph2car_grid(tt[r0:r1], tp[r0:r1])
    s3 = np.swapaxes(s3.reshape((3, -1)), 0, 1)
    t3 = np.swapaxes(t3.reshape((3, -1)), 0, 1)
    return np.dot(s3.T, t3)


def rotation_from_covariance(hmat):
    """pointcloud_utils.py:31-32."""
    u, _, vt = np.linalg.svd(hmat)
    return np.dot(vt.T, u.T)


def flow2rotation_3d(erp_flow):
    """flow_warp.py:91-133 (mask_method='center')."""
    return rotation_from_covariance(flow_cross_covariance(erp_flow))


def flow2rotation_2d(erp_flow):
    """flow_warp.py:136-187 with use_weight=False (use_weight=True raises numpy's AxisError at :183 in the
    reference: a 1-D array averaged over axis 1).  Rewrites the caller's u plane like the reference (:151)."""
    h, w = erp_flow.shape[0], erp_flow.shape[1]
    erp_flow = erp_of_wraparound(erp_flow)
    theta_delta_array = 2.0 * np.pi * (erp_flow[:, :, 0] / w)
    theta_delta = np.mean(theta_delta_array)
    delta = theta_delta / (2.0 * np.pi)
    col_start, col_end = int(w * (0.5 - delta)), int(w * (0.5 + delta))
    if delta < 0:
        col_start, col_end = col_end, col_start
    centre = np.zeros((h, w), dtype=bool)
    centre[:, col_start:col_end] = True
    flow_sign = np.sign(np.sum(np.sign(erp_flow[centre, 1])))
    if flow_sign < 0:
        chosen = np.logical_and(erp_flow[:, :, 1] < 0, centre)
    else:
        chosen = np.logical_and(erp_flow[:, :, 1] > 0, centre)
    with np.errstate(invalid="ignore", divide="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            phi_delta = np.mean(-np.pi * (erp_flow[chosen, 1] / h))
    return theta_delta, phi_delta


def rot_sph2mat(theta, phi, degrees_=True):
    """spherical_coordinates.py:221-224."""
    from scipy.spatial.transform import Rotation
    return Rotation.from_euler("xyz", [phi, theta, 0], degrees=degrees_).as_matrix()


def global_rotation_warping(erp_image, erp_flow, forward_warp=True, rotation_type="3D"):
    """flow_warp.py:190-229."""
    if rotation_type == "2D":
        theta_delta, phi_delta = flow2rotation_2d(erp_flow)
        if not forward_warp:
            theta_delta, phi_delta = -theta_delta, -phi_delta
        rotation_mat = rot_sph2mat(theta_delta, phi_delta, False)
    else:
        rotation_mat = flow2rotation_3d(erp_flow)
        if not forward_warp:
            rotation_mat = rotation_mat.T
    out = rotate_erp_array(erp_image, rotation_mat)
    if erp_image.dtype == np.uint8:
        out = out.astype(np.uint8)
    return out, rotation_mat


def flow_rotate_endpoint(optical_flow, rotation_matrix, wraparound=False):
    """projection.py:120-159 for a rotation matrix."""
    h, w = optical_flow.shape[:2]
    sx, sy = np.meshgrid(np.linspace(0, w, w, endpoint=False), np.linspace(0, h, h, endpoint=False))
    ex, ey = erp_pixles_modulo(sx + optical_flow[:, :, 0], sy + optical_flow[:, :, 1], w, h)
    field = rotation2erp_motion_vector((h, w), rotation_matrix, wraparound=True)
    ru = map_coordinates_order1(field[:, :, 0], ey, ex, "wrap")
    rv = map_coordinates_order1(field[:, :, 1], ey, ex, "wrap")
    ex, ey = erp_pixles_modulo(ex + ru, ey + rv, w, h)
    out = np.stack((ex - sx, ey - sy), axis=-1)
    if wraparound:
        out = erp_of_wraparound(out)
    return out


# ------------------------------------------------------------------------------------------------
# stand-alone face weights -- projection.py:15-117 (the stitch functions above restate them in fused form)
# ------------------------------------------------------------------------------------------------

def _warp_error(face_x, face_y, flow_uv, image_src, image_tar):
    n, channels = np.shape(face_x)[0], image_src.shape[2]
    diff = np.zeros((n, channels), dtype=np.float64)
    for ch in range(channels):
        t = map_coordinates_order1(image_tar[:, :, ch], flow_uv[:, 1], flow_uv[:, 0], "constant", 255)
        s = map_coordinates_order1(image_src[:, :, ch], face_y, face_x, "constant", 255)
        diff[:, ch] = np.absolute(np.asarray(t, dtype=np.float64) - np.asarray(s, dtype=np.float64))
    return diff


def get_blend_weight_ico(face_x, face_y, flow_uv, image_src, image_tar):
    """projection.py:15-61, weight_type 'image_warp_error'."""
    diff = _warp_error(face_x, face_y, flow_uv, image_src, image_tar)
    return np.exp(-(np.mean(diff, axis=1) / np.mean(diff)))


def get_blend_weight_cubemap(face_x, face_y, flow_uv, image_src, image_tar):
    """projection.py:64-117, weight_type 'image_warp_error'."""
    rgb_diff = np.linalg.norm(_warp_error(face_x, face_y, flow_uv, image_src, image_tar), axis=1)
    weights = np.ones(np.shape(face_x)[0], dtype=np.float64)
    nz = rgb_diff != 0.0
    weights[nz] = 0.95 / rgb_diff[nz]
    return weights


# ------------------------------------------------------------------------------------------------
# metrics -- flow_evaluate.py:23-217
# ------------------------------------------------------------------------------------------------

def available_pixel(flow, of_mask=None, unknown_value=1e7):
    """flow_evaluate.py:23-54, operator precedence of `index_unlarge` included."""
    u, v = flow[:, :, 0], flow[:, :, 1]
    with np.errstate(invalid="ignore"):
        unlarge = np.logical_not(abs(u) > unknown_value) | (abs(v) > unknown_value)
        unnan = np.logical_not(np.isnan(u) | np.isnan(v))
        unzero = (np.absolute(u) > 0) | (np.absolute(v) > 0)
    valid = unlarge & unnan & unzero
    if of_mask is not None:
        valid = valid & of_mask
    return valid


def great_circle_distance_uv(t1, p1, t2, p2):
    """spherical_coordinates.py:28-52."""
    a = np.sin((p2 - p1) * 0.5) ** 2 + np.cos(p1) * np.cos(p2) * np.sin((t2 - t1) * 0.5) ** 2
    return np.abs(1 * (2 * np.arctan2(np.sqrt(a), np.sqrt(1 - a))))


def epe_mat(gt, ev, spherical=False, of_mask=None):
    """flow_evaluate.py:70-104 (and RMSE_mat :118-153).  Works on copies; the reference zeroes its inputs in place."""
    gt, ev = np.array(gt, dtype=np.float64), np.array(ev, dtype=np.float64)
    ok = available_pixel(gt, of_mask)
    gt[~ok] = 0
    ev[~ok] = 0
    if spherical:
        h = gt.shape[0]
        g = flow_warp_meshgrid(gt[:, :, 0], gt[:, :, 1])
        e = flow_warp_meshgrid(ev[:, :, 0], ev[:, :, 1])
        gth, gph = erp2sph_grid(g[0], g[1], h)
        eth, eph = erp2sph_grid(e[0], e[1], h)
        return great_circle_distance_uv(gth, gph, eth, eph), ok
    return np.sqrt((gt[:, :, 0] - ev[:, :, 0]) ** 2 + (gt[:, :, 1] - ev[:, :, 1]) ** 2), ok


def aae_mat(gt, ev, of_mask=None):
    """flow_evaluate.py:167-217, planar branch."""
    gt, ev = np.array(gt, dtype=np.float64), np.array(ev, dtype=np.float64)
    ok = available_pixel(gt, of_mask)
    gt[~ok] = 0
    ev[~ok] = 0
    cross = gt[:, :, 0] * ev[:, :, 1] - gt[:, :, 1] * ev[:, :, 0]
    dot = gt[:, :, 0] * ev[:, :, 0] + gt[:, :, 1] * ev[:, :, 1]
    return abs(np.arctan2(cross, dot)), ok


def flow_metrics(gt, ev, spherical=False, of_mask=None):
    """EPE, RMSE, AAE as flow_evaluate.py:57-67,107-115,156-164 compute them (AAE: always planar, unmasked)."""
    e, ok = epe_mat(gt, ev, spherical, of_mask)
    a, ok_a = aae_mat(gt, ev, None)
    return np.sum(e) / np.sum(ok), np.sqrt(np.sum(np.square(e)) / np.sum(ok)), np.sum(a) / np.sum(ok_a)


def opticalflow_metric(flo_gt, flo_eva, flo_mask=None):
    """flow_evaluate.py:220-246 without the error-map images: AAE, EPE, RMS, then the same three "spherical" (AAE is
    always planar and unmasked, :156-164).  Every call zeroes the unusable pixels of BOTH arguments in place
    (:84-85, :132-133, :185-186), so the later calls see what the earlier ones left -- reproduced here."""
    def zero(mask):
        ok = available_pixel(flo_gt, mask)
        flo_gt[~ok] = 0
        flo_eva[~ok] = 0

    out = []
    for spherical in (False, True):
        a, ok_a = aae_mat(flo_gt, flo_eva, None)
        zero(None)
        out.append(np.sum(a) / np.sum(ok_a))
        e, ok = epe_mat(flo_gt, flo_eva, spherical, flo_mask)
        zero(flo_mask)
        out.append(np.sum(e) / np.sum(ok))
        e, ok = epe_mat(flo_gt, flo_eva, spherical, flo_mask)
        zero(flo_mask)
        out.append(np.sqrt(np.sum(np.square(e)) / np.sum(ok)))
    return tuple(out)


# ------------------------------------------------------------------------------------------------
# colour-wheel visualisation (row f4) -- flow_vis.py:19-201, image_evaluate.py:8-32
# ------------------------------------------------------------------------------------------------

def make_colorwheel():
    """flow_vis.py:19-66: the 55 x 3 Middlebury wheel."""
    ry, yg, gc, cb, bm, mr = 15, 6, 4, 11, 13, 6
    wheel = np.zeros((ry + yg + gc + cb + bm + mr, 3))
    col = 0
    wheel[0:ry, 0] = 255
    wheel[0:ry, 1] = np.floor(255 * np.arange(0, ry) / ry)
    col += ry
    wheel[col:col + yg, 0] = 255 - np.floor(255 * np.arange(0, yg) / yg)
    wheel[col:col + yg, 1] = 255
    col += yg
    wheel[col:col + gc, 1] = 255
    wheel[col:col + gc, 2] = np.floor(255 * np.arange(0, gc) / gc)
    col += gc
    wheel[col:col + cb, 1] = 255 - np.floor(255 * np.arange(cb) / cb)
    wheel[col:col + cb, 2] = 255
    col += cb
    wheel[col:col + bm, 2] = 255
    wheel[col:col + bm, 0] = np.floor(255 * np.arange(0, bm) / bm)
    col += bm
    wheel[col:col + mr, 2] = 255 - np.floor(255 * np.arange(mr) / mr)
    wheel[col:col + mr, 0] = 255
    return wheel


def flow_uv_to_colors(u, v, convert_to_bgr=False):
    """flow_vis.py:77-121."""
    image = np.zeros((u.shape[0], u.shape[1], 3), np.uint8)
    wheel = make_colorwheel()
    rows = wheel.shape[0]
    rad = np.sqrt(np.square(u) + np.square(v))
    angle = np.arctan2(-v, -u) / np.pi
    idx = (angle + 1) / 2 * (rows - 1)
    k0 = np.floor(idx).astype(np.int32)
    k1 = k0 + 1
    k1[k1 == rows] = 0
    ratio = idx - k0
    for i in range(3):
        c0 = wheel[:, i][k0] / 255.0
        c1 = wheel[:, i][k1] / 255.0
        color = (1 - ratio) * c0 + ratio * c1
        inside = rad <= 1
        color[inside] = 1 - rad[inside] * (1 - color[inside])
        color[~inside] = color[~inside] * 0.75
        image[:, :, 2 - i if convert_to_bgr else i] = np.floor(255 * color)
    return image


def flow_to_color(flow_uv, clip_flow=None, convert_to_bgr=False, min_ratio=0.0, max_ratio=1.0):
    """flow_vis.py:124-163 (without the matplotlib bar)."""
    if min_ratio != 0 and max_ratio != 1.0:
        flat = flow_uv.flatten()
        lo, hi = int(flat.size * min_ratio), int((flat.size - 1) * max_ratio)
        clip_flow = (np.partition(flat, lo)[lo], np.partition(flat, hi)[hi])
    if clip_flow is not None:
        flow_uv = np.clip(flow_uv, clip_flow[0], clip_flow[1])
    u, v = flow_uv[:, :, 0], flow_uv[:, :, 1]
    rad_max = np.max(np.sqrt(np.square(u) + np.square(v)))
    return flow_uv_to_colors(u / (rad_max + 1e-5), v / (rad_max + 1e-5), convert_to_bgr)
