"""Host-side gnomonic helpers with the reference's names and argument meaning (src/gnomonic_projection.py).

Used on the host to build the face table (three vertices per face).  The per-pixel forward projection
is in k_plan_stage2 (csrc/tif_plan.cu), the inverse in k_plan_stage1 and lift_endpoint (csrc/tif_exec.cu).
"""
import numpy as np

try:
    from . import spherical_coordinates
except ImportError:      # imported by bare name (drop-in mode)
    import spherical_coordinates


def gnomonic_projection(theta, phi, theta_0, phi_0, label_outrange_pixel=False):
    """(theta, phi) -> tangent-plane (x, y) (src/gnomonic_projection.py:18-56).  Third value: mask of points
    on the far hemisphere when label_outrange_pixel, else None."""
    cos_c = np.sin(phi_0) * np.sin(phi) + np.cos(phi_0) * np.cos(phi) * np.cos(theta - theta_0)
    degenerate = cos_c == 0
    if np.any(degenerate):
        cos_c = np.where(degenerate, np.finfo(float).eps, cos_c)
    x = np.cos(phi) * np.sin(theta - theta_0) / cos_c
    y = (np.cos(phi_0) * np.sin(phi) - np.sin(phi_0) * np.cos(phi) * np.cos(theta - theta_0)) / cos_c
    if np.any(degenerate):
        x = np.where(degenerate, 0, x)
        y = np.where(degenerate, 0, y)
    if label_outrange_pixel:
        far = spherical_coordinates.great_circle_distance_uv(theta, phi, theta_0, phi_0) >= (np.pi * 0.5)
        return x, y, far
    return x, y, None


def reverse_gnomonic_projection(x, y, theta_0, phi_0, hemisphere_index=None):
    """Tangent-plane (x, y) -> (theta, phi), theta continuous (src/gnomonic_projection.py:59-95)."""
    rho = np.sqrt(x ** 2 + y ** 2)
    degenerate = rho == 0
    if np.any(degenerate):
        rho = np.where(degenerate, np.finfo(float).eps, rho)
    c = np.arctan2(rho, 1)
    phi_ = np.arcsin(np.cos(c) * np.sin(phi_0) + (y * np.sin(c) * np.cos(phi_0)) / rho)
    theta_ = theta_0 + np.arctan2(x * np.sin(c), rho * np.cos(phi_0) * np.cos(c) - y * np.sin(phi_0) * np.sin(c))
    if hemisphere_index is not None:
        far = ~hemisphere_index
        t2, p2 = spherical_coordinates.sph_coord_modulo(theta_ + np.pi, -phi_)
        theta_ = np.where(far, t2, theta_)
        phi_ = np.where(far, p2, phi_)
    if np.any(degenerate):
        phi_ = np.where(degenerate, phi_0, phi_)
        theta_ = np.where(degenerate, theta_0, theta_)
    return theta_, phi_


def _range_or_unit(coord_gnom_xy_range):
    if coord_gnom_xy_range is None:
        return -1.0, 1.0, -1.0, 1.0
    return tuple(coord_gnom_xy_range[:4])


def gnomonic2pixel(coord_gnom_x, coord_gnom_y, padding_size, tangent_image_width, tangent_image_height=None,
                   coord_gnom_xy_range=None):
    """Gnomonic (x right, y up) -> integer pixel (x right, y down), nearest (src/gnomonic_projection.py:98-149)."""
    if tangent_image_height is None:
        tangent_image_height = tangent_image_width
    x_min, x_max, y_min, y_max = _range_or_unit(coord_gnom_xy_range)
    ratio_w = (tangent_image_width - 1.0) / (x_max - x_min + padding_size * 2.0)
    ratio_h = (tangent_image_height - 1.0) / (y_max - y_min + padding_size * 2.0)
    px = ((coord_gnom_x - x_min + padding_size) * ratio_w + 0.5)
    py = (-(coord_gnom_y - y_max - padding_size) * ratio_h + 0.5)
    return np.asarray(px).astype(int), np.asarray(py).astype(int)


def pixel2gnomonic(coord_pixel_x, coord_pixel_y, padding_size, tangent_image_width, tangent_image_height=None,
                   coord_gnom_xy_range=None):
    """Float pixel -> gnomonic coordinate (src/gnomonic_projection.py:152-196)."""
    if tangent_image_height is None:
        tangent_image_height = tangent_image_width
    x_min, x_max, y_min, y_max = _range_or_unit(coord_gnom_xy_range)
    ratio_w = (tangent_image_width - 1.0) / (abs(x_max - x_min) + padding_size * 2.0)
    ratio_h = (tangent_image_height - 1.0) / (abs(y_max - y_min) + padding_size * 2.0)
    return coord_pixel_x / ratio_w + x_min - padding_size, - coord_pixel_y / ratio_h + y_max + padding_size
