"""Generate the golden fixtures in tests/golden/ from the reference checkout.

Run HERE (container with /root/reference): python tests/golden/make_golden.py
Only the reference's pure-Python modules are imported (chroma.make, chroma.tools,
chroma.geometry, chroma.detector, chroma.sample); its GPU host code needs PyCUDA
and is not importable.  Outputs (committed):
  cube_rays.npz     reference make.cube(1000) mesh + a subsample of the rays /
                    distances of the reference's own golden vector
                    test/data/ray_intersection.npy (rays start at the film pixel,
                    SURVEY.md section 4)
  flatten.npz       Geometry/Detector.flatten() of a small scene built with the
                    reference's classes (vertices, triangles, solid_id, indices)
  sphere_mesh.npz   reference make.sphere(1000, 16) (mesh-builder cross-check)
  host_helpers.npz  chroma.transform (rotation matrices and rotated points, gen_rot, get_perp), triangle /
                    vertex counts, area and signed volume of chroma.make's other builders, and
                    chroma.sample.flashlight under a fixed NumPy seed
"""
import os
import sys
import numpy as np

REF = os.environ.get('CHROMA_REFERENCE', '/root/reference')
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))

import chroma.make as rmake            # noqa: E402
import chroma.tools as rtools          # noqa: E402
import chroma.geometry as rgeo         # noqa: E402
import chroma.detector as rdet         # noqa: E402
from chroma.transform import normalize  # noqa: E402


def film_rays(size=(800, 600), width=35.0, focal_length=18.0):
    """Ray set that reproduces test/data/ray_intersection.npy: the grid of
    chroma/tools.py:207-239 with rays STARTING at the film pixel."""
    axis1, axis2 = np.array((0, 0, 1.0)), np.array((1.0, 0, 0))
    height = width * (size[1] / float(size[0]))
    dx0, dx1 = width / size[0], height / size[1]
    yy, xx = np.meshgrid(np.arange(size[1]), np.arange(size[0]))
    n = size[0] * size[1]
    grid = -np.tile(axis2, (n, 1)) * xx.ravel()[:, None] * dx0 + np.tile(axis1, (n, 1)) * yy.ravel()[:, None] * dx1
    grid += axis2 * width / 2 - axis1 * height / 2
    grid -= np.cross(axis1, axis2) * focal_length
    return grid, normalize(-grid)


def main():
    cube = rmake.cube(size=1000.0)
    pos, dirs = film_rays()
    # sanity: same directions as the reference's from_film()
    _, dref = rtools.from_film()
    assert np.allclose(dirs, dref)
    gold = np.load(os.path.join(REF, 'test', 'data', 'ray_intersection.npy'))
    assert gold.shape == (480000,)
    sel = np.arange(0, 480000, 13)
    zeros = np.flatnonzero(gold == 0)            # the 56 diagonal rays the old triangle test missed
    sel = np.unique(np.concatenate([sel, zeros]))
    np.savez_compressed(os.path.join(HERE, 'cube_rays.npz'),
                        vertices=cube.vertices.astype(np.float32), triangles=cube.triangles.astype(np.uint32),
                        index=sel.astype(np.int32), pos=pos[sel].astype(np.float32), dir=dirs[sel].astype(np.float32),
                        distance=gold[sel].astype(np.float32))

    sph = rmake.sphere(1000.0, 16)
    np.savez_compressed(os.path.join(HERE, 'sphere_mesh.npz'), vertices=sph.vertices, triangles=sph.triangles)

    # small detector: shell + 3 "PMTs" (cubes), rotations and displacements
    m_a, m_b = rgeo.Material('a'), rgeo.Material('b')
    s_x = rgeo.Surface('x')
    det = rdet.Detector(m_a)
    det.add_solid(rgeo.Solid(rmake.sphere(500.0, 8), m_a, m_a, surface=s_x))
    rot = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]], dtype=np.float32)
    small = rmake.cube(10.0)
    for k in range(3):
        det.add_pmt(rgeo.Solid(small, m_b, m_a), rotation=rot if k % 2 else None, displacement=(100.0 * k, 5.0, -20.0 * k))
    det.flatten()
    np.savez_compressed(os.path.join(HERE, 'flatten.npz'), vertices=det.mesh.vertices, triangles=det.mesh.triangles,
                        solid_id=det.solid_id, colors=det.colors,
                        material1_is_b=np.array([det.unique_materials[i] is m_b for i in det.material1_index]),
                        material2_is_b=np.array([det.unique_materials[i] is m_b for i in det.material2_index]),
                        surface_index=det.surface_index, surface_is_x=np.array([i >= 0 and det.unique_surfaces[i] is s_x for i in det.surface_index]),
                        solid_id_to_channel_index=det.solid_id_to_channel_index,
                        small_vertices=small.vertices, small_triangles=small.triangles,
                        shell_vertices=det.solids[0].mesh.vertices, shell_triangles=det.solids[0].mesh.triangles)
    helpers(os.path.join(HERE, 'host_helpers.npz'))
    print('wrote fixtures to', HERE)


HELPER_AXIS, HELPER_PHI = np.array([0.3, -0.5, 0.8]), 0.7
GEN_ROT_PAIRS = (([1, 0, 0], [0, 1, 0]), ([1, 0, 0], [1, 0, 0]), ([0, 1, 0], [0, 1, 0]), ([0, 0, 1], [0, 0, -1]),
                 ([1, 2, 3], [-2, 0.5, 1]))


def builder_cases():
    ang = np.linspace(0, 2 * np.pi, 6, endpoint=False)
    hexagon = (np.cos(ang), np.sin(ang))
    return (('linear_extrude', hexagon + (2.0,), {}),
            ('linear_extrude', hexagon + (2.0,), dict(x2=0.5 * hexagon[0], y2=0.5 * hexagon[1], center=(1, 2, 3))),
            ('linear_extrude', hexagon + (2.0,), dict(endcaps=False)),
            ('cylinder_along_z', (10.0, 30.0, 20), {}), ('segmented_cylinder', (10.0, 30.0, 16, 40), {}),
            ('torus', (2.0, 10.0, 16, 12), {}), ('convex_polygon', hexagon, {}), ('cylinder', (5.0, 8.0, 3.0, 12), {}))


def mesh_stats(mesh):
    """(triangles, vertices, area, signed volume): independent of vertex numbering, sensitive to orientation."""
    v = np.asarray(mesh.vertices, dtype=np.float64)[np.asarray(mesh.triangles)]
    cr = np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0])
    return np.array([len(mesh.triangles), len(mesh.vertices), 0.5 * np.linalg.norm(cr, axis=1).sum(),
                     np.einsum('ij,ij->i', v[:, 0], cr).sum() / 6.0])


def helpers(path):
    import chroma.transform as rtr
    import chroma.sample as rsample
    x = np.random.default_rng(0).normal(size=(5, 3))
    out = {'points': x,
           'matrix': rtr.make_rotation_matrix(HELPER_PHI, HELPER_AXIS),
           'rotate': rtr.rotate(x, HELPER_PHI, HELPER_AXIS),
           'rotate_many': rtr.rotate(x, np.linspace(0, 1, 5), HELPER_AXIS),
           'rotate_matrix': rtr.rotate_matrix(x, HELPER_PHI, HELPER_AXIS),
           'get_perp': rtr.get_perp(HELPER_AXIS),
           'gen_rot': np.array([rtr.gen_rot(np.array(a, float), np.array(b, float)) for a, b in GEN_ROT_PAIRS]),
           'builders': np.array([mesh_stats(getattr(rmake, n)(*a, **k)) for n, a, k in builder_cases()])}
    np.random.seed(5)
    out['flashlight'] = rsample.flashlight(0.3, (1, 2, 3), 1000)
    np.random.seed(5)
    out['flashlight_one'] = rsample.flashlight()
    np.savez_compressed(path, **out)


if __name__ == '__main__':
    main()
