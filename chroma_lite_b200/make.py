"""Mesh primitives used as fixtures and by the demo detector (role of
chroma/make.py).  Written from scratch: a surface of revolution is built ring by
ring about the y axis; degenerate pole triangles and duplicate pole vertices are
removed by Mesh."""
import numpy as np

from .geometry import Mesh


def rotate_extrude(x, y, nsteps=64):
    """Revolve the profile (x_i, y_i) about the y axis in `nsteps` angular steps.
    A counter-clockwise profile (bottom to top with x >= 0) gives outward normals."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if len(x) != len(y):
        raise Exception('`x` and `y` arrays must have the same length.')
    npts = len(x)
    ang = np.linspace(0.0, 2.0 * np.pi, nsteps, endpoint=False)
    # vertex (ring j, profile point i) -> index j*npts + i
    vx = np.outer(np.cos(ang), x)
    vz = np.outer(np.sin(ang), x)
    vy = np.tile(y, (nsteps, 1))
    vertices = np.stack([vx, vy, vz], axis=-1).reshape(-1, 3)
    j = np.arange(nsteps)
    jn = (j + 1) % nsteps
    i = np.arange(npts - 1)
    J, I = np.meshgrid(j, i, indexing='ij')
    JN = jn[J]
    a = (J * npts + I).ravel()
    b = (J * npts + I + 1).ravel()
    c = (JN * npts + I + 1).ravel()
    d = (JN * npts + I).ravel()
    triangles = np.concatenate([np.stack([a, b, c], axis=1), np.stack([a, c, d], axis=1)])
    return Mesh(vertices, triangles, remove_duplicate_vertices=True)


def sphere(radius, nsteps=64):
    t = np.linspace(-np.pi / 2, np.pi / 2, nsteps)
    return rotate_extrude(radius * np.cos(t), radius * np.sin(t), nsteps)


def cylinder(radius, height, radius2=None, nsteps=64):
    if radius2 is None:
        radius2 = radius
    return rotate_extrude([0, radius, radius2, 0], [-height / 2.0, -height / 2.0, height / 2.0, height / 2.0], nsteps)


def box(dx, dy, dz, center=(0, 0, 0)):
    """Axis-aligned box of 12 triangles with outward normals."""
    hx, hy, hz = dx / 2.0, dy / 2.0, dz / 2.0
    v = np.array([[-hx, -hy, -hz], [hx, -hy, -hz], [hx, hy, -hz], [-hx, hy, -hz],
                  [-hx, -hy, hz], [hx, -hy, hz], [hx, hy, hz], [-hx, hy, hz]], dtype=np.float64)
    v += np.asarray(center, dtype=np.float64)
    t = np.array([[0, 2, 1], [0, 3, 2], [4, 5, 6], [4, 6, 7], [0, 1, 5], [0, 5, 4],
                  [1, 2, 6], [1, 6, 5], [2, 3, 7], [2, 7, 6], [3, 0, 4], [3, 4, 7]])
    return Mesh(v, t)


def cube(size, height=None, center=(0, 0, 0)):
    return box(size, size, size if height is None else height, center=center)


def subdivide(mesh, times=1):
    """1 -> 4 midpoint subdivision (used to grow the ray-microbench mesh)."""
    v = np.asarray(mesh.vertices, dtype=np.float64)
    t = np.asarray(mesh.triangles, dtype=np.int64)
    for _ in range(times):
        e = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]])
        e.sort(axis=1)
        uniq, inv = np.unique(e, axis=0, return_inverse=True)
        inv = np.asarray(inv).reshape(-1)
        mid = 0.5 * (v[uniq[:, 0]] + v[uniq[:, 1]])
        base = len(v)
        v = np.concatenate([v, mid])
        n = len(t)
        m01, m12, m20 = base + inv[:n], base + inv[n:2 * n], base + inv[2 * n:]
        t = np.concatenate([np.stack([t[:, 0], m01, m20], 1), np.stack([m01, t[:, 1], m12], 1),
                            np.stack([m20, m12, t[:, 2]], 1), np.stack([m01, m12, m20], 1)])
    return Mesh(v, t, remove_duplicate_vertices=False)
