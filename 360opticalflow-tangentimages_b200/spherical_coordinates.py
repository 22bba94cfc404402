"""Host-side spherical helpers with the reference's names and argument meaning (src/spherical_coordinates.py).

Used on the host for the face table and the per-row / per-column tables of the plan; the per-pixel
arithmetic of the hot path lives in csrc/tif_plan.cu and csrc/tif_exec.cu.  The rotation helpers of
the reference (rotate_erp_array, rotation2erp_motion_vector, rot_*) belong to the de-rotation stages
and are out of scope.
"""
import numpy as np


def great_circle_distance_uv(points_1_theta, points_1_phi, points_2_theta, points_2_phi, radius=1):
    """Haversine distance (src/spherical_coordinates.py:28-52)."""
    d_theta = points_2_theta - points_1_theta
    d_phi = points_2_phi - points_1_phi
    a = np.sin(d_phi * 0.5) ** 2 + np.cos(points_1_phi) * np.cos(points_2_phi) * np.sin(d_theta * 0.5) ** 2
    return np.abs(radius * (2 * np.arctan2(np.sqrt(a), np.sqrt(1 - a))))


def great_circle_distance(points_1, points_2, radius=1):
    """src/spherical_coordinates.py:12-25."""
    return great_circle_distance_uv(points_1[0], points_1[1], points_2[0], points_2[1], radius)


def erp_sph_modulo(theta, phi):
    """theta -> [-pi, pi), phi -> (-pi/2, pi/2] (src/spherical_coordinates.py:55-61)."""
    return np.remainder(theta + np.pi, 2 * np.pi) - np.pi, -(np.remainder(-phi + 0.5 * np.pi, np.pi) - 0.5 * np.pi)


def sph_coord_modulo(theta, phi):
    """theta -> [-pi, pi), phi -> [-pi/2, pi/2) (src/spherical_coordinates.py:234-239)."""
    return np.remainder(theta + np.pi, np.pi * 2.0) - np.pi, np.remainder(phi + 0.5 * np.pi, np.pi) - 0.5 * np.pi


def erp2sph(erp_points, erp_image_height=None, sph_modulo=False):
    """Pixel centre -> (theta, phi) (src/spherical_coordinates.py:64-101).  erp_points is (x, y) or [2, ...]."""
    if erp_image_height is None:
        height, width = np.shape(erp_points)[1], np.shape(erp_points)[2]
    else:
        height, width = erp_image_height, erp_image_height * 2
    theta = erp_points[0] * (2 * np.pi / width) + np.pi / width - np.pi
    phi = -(erp_points[1] * (np.pi / height) + np.pi / height * 0.5) + 0.5 * np.pi
    if sph_modulo:
        theta, phi = erp_sph_modulo(theta, phi)
    theta = np.where(theta == np.pi, -np.pi, theta)
    phi = np.where(phi == -0.5 * np.pi, 0.5 * np.pi, phi)
    return np.stack((theta, phi))


def sph2erp(theta, phi, erp_image_height, sph_modulo=False):
    """(theta, phi) -> float ERP pixel (src/spherical_coordinates.py:104-125)."""
    if sph_modulo:
        theta, phi = erp_sph_modulo(theta, phi)
    erp_image_width = 2 * erp_image_height
    erp_x = (theta + np.pi) / (2.0 * np.pi / erp_image_width) - 0.5
    erp_y = (-phi + 0.5 * np.pi) / (np.pi / erp_image_height) - 0.5
    return erp_x, erp_y


def car2sph(points_car, min_radius=1e-10):
    """[N,3] cartesian (+x right, +y down, +z forward) -> [N,2] (theta, phi) (src/spherical_coordinates.py:128-148)."""
    points_car = np.asarray(points_car, dtype=np.float64)
    radius = np.linalg.norm(points_car, axis=1)
    ok = radius > min_radius
    theta = np.zeros(points_car.shape[0], np.float64)
    phi = np.zeros(points_car.shape[0], np.float64)
    theta[ok] = np.arctan2(points_car[:, 0][ok], points_car[:, 2][ok])
    phi[ok] = -np.arcsin(np.divide(points_car[:, 1][ok], radius[ok]))
    return np.stack((theta, phi), axis=1)


def sph2car(theta, phi, radius=1.0):
    """(theta, phi) -> [3, ...] cartesian (src/spherical_coordinates.py:151-169)."""
    x = radius * np.cos(phi) * np.sin(theta)
    z = radius * np.cos(phi) * np.cos(theta)
    y = -radius * np.sin(phi)
    return np.stack((x, y, z), axis=0)


def sph_wraparound(src_theta, tar_theta):
    """Masks of lines that cross the +-pi seam the long way round (src/spherical_coordinates.py:242-258)."""
    long_line = np.abs(tar_theta - src_theta) > np.pi
    minus = long_line & (src_theta < 0) & (tar_theta > 0)
    plus = long_line & (src_theta > 0) & (tar_theta < 0)
    return minus, plus
