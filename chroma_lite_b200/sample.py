"""Random direction samplers (role of chroma/sample.py)."""
import numpy as np


def uniform_sphere(size=None, dtype=np.double, rng=None):
    """Isotropic unit vectors; uses numpy's global RNG unless `rng` is given."""
    r = np.random if rng is None else rng
    theta, u = r.uniform(0.0, 2 * np.pi, size), r.uniform(-1.0, 1.0, size)
    c = np.sqrt(1 - u ** 2)
    if size is None:
        return np.array([c * np.cos(theta), c * np.sin(theta), u], dtype=dtype)
    pts = np.empty((np.prod(size), 3), dtype=dtype)
    pts[:, 0], pts[:, 1], pts[:, 2] = c * np.cos(theta), c * np.sin(theta), u
    return pts


def flashlight(phi=np.pi / 4, direction=(0, 0, 1), size=None, dtype=np.double, rng=None):
    """Unit vectors uniform in the cone of half-angle `phi` about `direction`
    (chroma/sample.py:32-56)."""
    from .transform import rotate
    r = np.random if rng is None else rng
    theta, u = r.uniform(0.0, 2 * np.pi, size), r.uniform(np.cos(phi), 1, size)
    c = np.sqrt(1 - u ** 2)
    if np.equal(direction, (0, 0, 1)).all():
        axis, angle = (0, 0, 1), 0.0
    else:
        axis = np.cross((0, 0, 1), direction)
        angle = -np.arccos(np.dot(direction, (0, 0, 1)) / np.linalg.norm(direction))
    if size is None:
        return rotate(np.array([c * np.cos(theta), c * np.sin(theta), u]), angle, axis)
    pts = np.empty((size, 3), dtype)
    pts[:, 0], pts[:, 1], pts[:, 2] = c * np.cos(theta), c * np.sin(theta), u
    return rotate(pts, angle, axis)


def normalize(x):
    x = np.asarray(x, dtype=np.double)
    return x / np.sqrt((x ** 2).sum(axis=-1))[..., None] if x.ndim > 1 else x / np.sqrt((x ** 2).sum())


def make_rotation_matrix(phi, n):
    """Rotation by phi about axis n (right-handed)."""
    n = normalize(n)
    K = np.array([[0, -n[2], n[1]], [n[2], 0, -n[0]], [-n[1], n[0], 0]])
    return np.identity(3) + np.sin(phi) * K + (1 - np.cos(phi)) * (K @ K)
