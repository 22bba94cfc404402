"""Host-side polygon helpers with the reference's names and argument meaning (src/polygon.py).

They only run on the host while the face table is built (3 vertices per face); the per-pixel
point-in-polygon test of the hot path is `pip_accept` in csrc/tif_plan.cu.
"""
import numpy as np


def find_intersection(p1, p2, p3, p4):
    """Intersection of the lines p1->p2 and p3->p4 (src/polygon.py:8-40); None when parallel."""
    d12 = (p2[0] - p1[0], p2[1] - p1[1])
    d34 = (p4[0] - p3[0], p4[1] - p3[1])
    denominator = (d12[1] * d34[0] - d12[0] * d34[1])
    if denominator == 0:
        return None
    t1 = ((p1[0] - p3[0]) * d34[1] + (p3[1] - p1[1]) * d34[0]) / denominator
    return [p1[0] + d12[0] * t1, p1[1] + d12[1] * t1]


def is_clockwise(point_list):
    """Shoelace sign test (src/polygon.py:43-56)."""
    total = 0
    n = len(point_list)
    for i in range(n):
        cur, nxt = point_list[i], point_list[(i + 1) % n]
        total += (nxt[0] - cur[0]) * (nxt[1] + cur[1])
    return total > 0


def inside_polygon_2d(points_list, polygon_points, on_line=False, eps=1e-4):
    """Ray-cast point-in-polygon with eps-wide edges (src/polygon.py:59-118).

    points_list [N,2], polygon_points [M,2] clockwise.  Returns bool[N].
    """
    points_list = np.asarray(points_list)
    px, py = points_list[:, 0], points_list[:, 1]
    inside = np.zeros(px.shape[0], dtype=bool)
    online = np.zeros(px.shape[0], dtype=bool)
    m = len(polygon_points)
    for k in range(m):
        x1, y1 = polygon_points[k][0], polygon_points[k][1]
        x2, y2 = polygon_points[(k + 1) % m][0], polygon_points[(k + 1) % m][1]
        hit = (py >= min(y1, y2)) & (py <= max(y1, y2)) & (px <= max(x1, x2))
        if abs(y1 - y2) <= eps:
            hit &= px >= min(x1, x2)
            cross_x = px
        else:
            cross_x = (py - y1) * (x2 - x1) / (y2 - y1) + x1
        online |= hit & (np.abs(px - cross_x) <= eps)
        inside ^= hit & (px <= cross_x)
    if on_line:
        return inside | online
    return inside & ~online


def enlarge_polygon(old_points, offset):
    """Offset every edge outward by `offset` (absolute units) and intersect neighbours (src/polygon.py:121-167)."""
    n = len(old_points)

    def shifted_edge(a, b):
        v = np.array([b[0] - a[0], b[1] - a[1]], float)
        v = v / np.linalg.norm(v) * offset
        normal = [-v[1], v[0]]
        return [a[0] + normal[0], a[1] + normal[1]], [b[0] + normal[0], b[1] + normal[1]]

    enlarged = []
    for j in range(n):
        i, k = (j - 1) % n, (j + 1) % n
        e1 = shifted_edge(old_points[i], old_points[j])
        e2 = shifted_edge(old_points[j], old_points[k])
        enlarged.append(find_intersection(e1[0], e1[1], e2[0], e2[1]))
    return enlarged
