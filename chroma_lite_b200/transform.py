"""Rotation helpers with the conventions of chroma/transform.py:1-60.

Mind the sense: the reference's ``make_rotation_matrix(phi, n)`` and ``rotate(x, phi, n)`` turn
points counter-clockwise "when looking towards +infinity" along ``n``, which is the transpose of the
right-handed matrix (``sample.make_rotation_matrix`` of this package).  ``Geometry.add_solid`` applies
a rotation as ``np.inner(vertices, rotation)`` like the reference, so matrices made here place solids
exactly as matrices made by the reference do.
"""
import numpy as np


def normalize(x):
    """x / |x| along the last axis (chroma/transform.py normalize)."""
    x = np.asarray(x, dtype=np.double)
    return x / np.sqrt((x ** 2).sum(axis=-1))[..., None] if x.ndim > 1 else x / np.sqrt((x ** 2).sum())


def make_rotation_matrix(phi, n):
    """Matrix M such that ``np.inner(x, M)`` == ``rotate(x, phi, n)`` (chroma/transform.py:29-41)."""
    n = normalize(n)
    c, s = np.cos(phi), np.sin(phi)
    cross = np.array([[0.0, n[2], -n[1]], [-n[2], 0.0, n[0]], [n[1], -n[0], 0.0]])
    return c * np.identity(3) + (1.0 - c) * np.outer(n, n) + s * cross


def rotate(x, phi, n):
    """Rotate points ``x`` by ``phi`` (scalar or one angle per point) about ``n``
    (chroma/transform.py:43-52)."""
    n = normalize(n)
    x = np.atleast_2d(x)
    phi = np.atleast_1d(phi)
    c, s = np.cos(phi)[:, None], np.sin(phi)[:, None]
    return (x * c + n * np.dot(x, n)[:, None] * (1.0 - c) + np.cross(x, n) * s).squeeze()


def rotate_matrix(x, phi, n):
    """rotate() through the matrix; a single angle only (chroma/transform.py:54-60)."""
    return np.inner(np.asarray(x), make_rotation_matrix(phi, n))


def get_perp(x):
    """Some vector perpendicular to ``x`` (chroma/transform.py:22-27)."""
    a = np.zeros(3)
    a[np.argmin(np.abs(x))] = 1.0
    return np.cross(a, x)


def gen_rot(a, b):
    """Matrix that takes direction ``a`` to ``-b`` (chroma/transform.py:3-19)."""
    a = np.asarray(a, dtype=np.double) / np.linalg.norm(a)
    b = np.asarray(b, dtype=np.double) / np.linalg.norm(b)
    if (a == -b).all():
        return np.identity(3)
    if (a == b).all():
        v = np.cross(a, [0.0, 1.0, 0.0]) if (a[1] == 0 and a[2] == 0) else np.cross(a, [1.0, 0.0, 0.0])
        return make_rotation_matrix(np.pi, v)
    return make_rotation_matrix(np.arccos(-np.dot(a, b)), np.cross(a, b))
