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


def normalize(x):
    x = np.asarray(x, dtype=np.double)
    return x / np.sqrt((x ** 2).sum(axis=-1))[..., None] if x.ndim > 1 else x / np.sqrt((x ** 2).sum())


def make_rotation_matrix(phi, n):
    """Rotation by phi about axis n (right-handed)."""
    n = normalize(n)
    K = np.array([[0, -n[2], n[1]], [n[2], 0, -n[0]], [-n[1], n[0], 0]])
    return np.identity(3) + np.sin(phi) * K + (1 - np.cos(phi)) * (K @ K)
