"""Host-side spherical helpers with the reference's names and argument meaning (src/spherical_coordinates.py).

Used on the host for the face table and the per-row / per-column tables of the plan; the per-pixel
arithmetic of the hot path lives in csrc/tif_plan.cu and csrc/tif_exec.cu.  The rotation helpers
(rotate_erp_array, rotation2erp_motion_vector: row f2 of SURVEY.md section 8) run on the device through
csrc/tif_rotate.cu; rot_sph2mat / rot_mat2sph are 3x3 host conversions.
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


def rotation2erp_motion_vector(array_size, rotation_matrix=None, wraparound=False):
    """ERP flow [H,W,2] float64 of a rigid rotation of the sphere (src/spherical_coordinates.py:185-218); computed
    per pixel on the device (k_rotation_mv)."""
    import torch
    from . import _lib, _ops
    if not torch.cuda.is_available():
        raise _lib.TifError("no CUDA device: rotation2erp_motion_vector has no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device())
    return _ops.rotation_motion_vector(rotation_matrix, array_size[0], array_size[1], wraparound, dev).cpu().numpy()


def rotate_erp_array(erp_image, rotation_mat=None):
    """Rotate an ERP image (src/spherical_coordinates.py:172-182): backward warp by the motion vector of R^T; both
    steps stay on the device, the float64 flow keeps scipy's uint8 rounding exact."""
    import torch
    from . import _lib, _ops
    if not torch.cuda.is_available():
        raise _lib.TifError("no CUDA device: rotate_erp_array has no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device())
    erp_image = np.asarray(erp_image)
    h, w = erp_image.shape[:2]
    flow = _ops.rotation_motion_vector(np.asarray(rotation_mat).T, h, w, False, dev)
    if erp_image.ndim == 3:
        mode, img = _lib.WARP_WRAP, erp_image
    else:
        mode, img = _lib.WARP_CONSTANT255, erp_image[:, :, None]
    if img.dtype == np.uint8:
        img_dev = torch.from_numpy(np.ascontiguousarray(img)).to(dev)[None]
    else:
        img_dev = torch.from_numpy(np.ascontiguousarray(img, dtype=np.float32)).to(dev)[None]
    out = _ops.ops.warp_backward(img_dev, flow[None], mode)[0].cpu().numpy().astype(erp_image.dtype, copy=False)
    return out if erp_image.ndim == 3 else out[:, :, 0]


def rot_sph2mat(theta, phi, degrees_=True):
    """src/spherical_coordinates.py:221-224."""
    from scipy.spatial.transform import Rotation
    return Rotation.from_euler("xyz", [phi, theta, 0], degrees=degrees_).as_matrix()


def rot_mat2sph(rot_mat, degrees_=True):
    """src/spherical_coordinates.py:227-231."""
    from scipy.spatial.transform import Rotation
    return tuple(Rotation.from_matrix(rot_mat).as_euler("xyz", degrees=degrees_)[:2])
