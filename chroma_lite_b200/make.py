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
    vz = -np.outer(np.sin(ang), x)          # the reference turns the profile about -y (chroma/make.py rotate_extrude)
    vy = np.tile(y, (nsteps, 1))
    vertices = np.stack([vx, vy, vz], axis=-1).reshape(-1, 3)
    # Quads between profile points i+1 -> i and angular steps j -> j+1, listed from the top of the
    # profile down, all steps of a profile segment together; first the triangles (upper, lower,
    # lower-next) of every quad, then (upper, lower-next, upper-next): the triangle ORDER of the
    # reference's builder (chroma/make.py mesh_grid), so that triangle ids of a flattened detector
    # are the reference's.
    upper, step = np.meshgrid(np.arange(npts - 1, 0, -1), np.arange(nsteps), indexing='ij')
    nxt = (step + 1) % nsteps
    hi, lo = (step * npts + upper).ravel(), (step * npts + upper - 1).ravel()
    hi_next, lo_next = (nxt * npts + upper).ravel(), (nxt * npts + upper - 1).ravel()
    triangles = np.concatenate([np.stack([hi, lo, lo_next], axis=1), np.stack([hi, lo_next, hi_next], axis=1)])
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


def mesh_grid(grid):
    """Triangle index rows for a grid of vertex indices that is periodic along its second axis:
    each cell (i, j) between rows i, i+1 and columns j, j+1 gives two triangles
    (index arithmetic of chroma/make.py:6-20)."""
    grid = np.asarray(grid)
    a, b = grid[:-1], grid[1:]
    an, bn = np.roll(a, -1, axis=1), np.roll(b, -1, axis=1)
    first = np.stack([a.ravel(), b.ravel(), bn.ravel()], axis=1)
    second = np.stack([a.ravel(), bn.ravel(), an.ravel()], axis=1)
    return np.concatenate([first, second])


def linear_extrude(x1, y1, height, x2=None, y2=None, center=None, endcaps=True):
    """Prism from the counter-clockwise polygon (x1, y1) at z = -height/2 to (x2, y2) (default: the
    same polygon) at z = +height/2; end caps are fans about the z axis (chroma/make.py:22-69)."""
    x1, y1 = np.asarray(x1, dtype=np.float64), np.asarray(y1, dtype=np.float64)
    if len(x1) != len(y1):
        raise Exception('`x` and `y` arrays must have the same length.')
    x2 = x1 if x2 is None else np.asarray(x2, dtype=np.float64)
    y2 = y1 if y2 is None else np.asarray(y2, dtype=np.float64)
    if len(x2) != len(y2) or len(x2) != len(x1):
        raise Exception('`x` and `y` arrays must have the same length.')
    n = len(x1)
    zero = np.zeros(n)
    rings = [np.column_stack([x1, y1, np.full(n, -height / 2.0)]), np.column_stack([x2, y2, np.full(n, height / 2.0)])]
    if endcaps:
        rings = ([np.column_stack([zero, zero, np.full(n, -height / 2.0)])] + rings +
                 [np.column_stack([zero, zero, np.full(n, height / 2.0)])])
    vertices = np.concatenate(rings)                       # ring r, corner j -> index r*n + j
    if center is not None:
        vertices = vertices + np.asarray(center, dtype=np.float64)
    # rows from the top ring down so that a counter-clockwise polygon gives outward normals
    grid = np.arange(len(vertices)).reshape(len(rings), n)[::-1]
    return Mesh(vertices, mesh_grid(grid), remove_duplicate_vertices=True)


def cylinder_along_z(radius, height, points=100):
    """Cylinder with its axis along z (chroma/make.py:106-108)."""
    angles = np.linspace(0.0, 2.0 * np.pi, points, endpoint=False)
    return linear_extrude(radius * np.cos(angles), radius * np.sin(angles), height)


def segmented_cylinder(radius, height, nsteps=64, nsegments=100):
    """Cylinder about y whose profile is cut into about `nsegments` pieces (chroma/make.py:121-129)."""
    nr = int((nsegments * radius / (2 * radius + height)) / 2)
    nh = int((nsegments * height / (2 * radius + height)) / 2)
    x = np.concatenate([np.linspace(0, radius, nr, endpoint=False), [radius] * nh,
                        np.linspace(radius, 0, nr, endpoint=False), [0]])
    y = np.concatenate([[-height / 2.0] * nr, np.linspace(-height / 2.0, height / 2.0, nh, endpoint=False),
                        [height / 2.0] * (nr + 1)])
    return rotate_extrude(x, y, nsteps)


def torus(radius, offset, nsteps=64, circle_steps=None):
    """Torus about y: barrel of `radius` centred `offset` from the axis (chroma/make.py:136-147)."""
    if circle_steps is None:
        circle_steps = nsteps
    t = np.linspace(0.0, 2.0 * np.pi, circle_steps)
    return rotate_extrude(radius * np.cos(t) + offset, radius * np.sin(t), nsteps)


def convex_polygon(x, y):
    """Fan triangulation of a convex polygon in the z = 0 plane, points given in order
    (chroma/make.py:149-163)."""
    x = np.asarray(x, dtype=np.float64)
    vertices = np.column_stack([x, np.asarray(y, dtype=np.float64), np.zeros_like(x)])
    k = np.arange(1, len(vertices) - 1)
    return Mesh(vertices, np.column_stack([np.zeros_like(k), k, k + 1]))


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
