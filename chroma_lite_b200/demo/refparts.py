"""The reference's own demo-detector building blocks, restored from a fixture.

BASELINE config 3 is quoted on the detector of chroma/demo/__init__.py:32-64 built from
chroma.demo.pmt.build_8inch_pmt (SNO PMT profile, chroma/demo/pmt.py:7-15, chroma/pmt.py:40-73)
and the tables of chroma/demo/optics.py (WCSim water, SNO glass, R7081HQE photocathode).  The
reference's modules cannot be imported where the engine runs, so their OUTPUT -- the PMT solids
as triangle arrays and the (wavelength, value) tables -- is kept in
tests/golden/ref_detector_parts.npz (written by tests/golden/make_ref_detector_parts.py, which
imports the reference) and turned back into this package's Solid / Material / Surface objects
here.  The placement (spiral, orientation, liner, DAQ response) is this package's own code and is
pinned against the reference's result by tests/test_ref_detector_cpu.py.
"""
import os
from math import sin, cos, sqrt

import numpy as np

from ..geometry import Mesh, Solid, Material, Surface
from ..detector import Detector
from ..make import sphere
from ..transform import make_rotation_matrix, normalize

DEFAULT_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
                            'tests', 'golden', 'ref_detector_parts.npz')


class Parts(object):
    """Materials, surfaces and solids of the fixture as this package's objects (one object per
    name, so that solids share them and Geometry.flatten() counts each once)."""

    def __init__(self, path=None):
        path = path or os.environ.get('CHROMA_B200_REF_PARTS') or DEFAULT_PATH
        if not os.path.exists(path):
            raise IOError('reference detector parts not found: %s (python tests/golden/make_ref_detector_parts.py '
                          'writes it where /root/reference is present)' % path)
        self.z = np.load(path)
        self.materials = [self._material(str(n)) for n in self.z['materials']]
        self.surfaces = [self._surface(str(n)) for n in self.z['surfaces']]

    def _table_object(self, obj, prefix):
        for key in self.z.files:
            if key.startswith(prefix):
                name = key[len(prefix):]
                value = self.z[key]
                obj.__dict__[name] = value if value.ndim else value.item()
        return obj

    def _material(self, name):
        return self._table_object(Material(name), 'mat.%s.' % name)

    def _surface(self, name):
        s = self._table_object(Surface(name), 'surf.%s.' % name)
        s.model, s.transmissive = int(s.model), int(s.transmissive)
        return s

    def material(self, name):
        return next(m for m in self.materials if m.name == name)

    def surface(self, name):
        return next(s for s in self.surfaces if s.name == name)

    def solid(self, name):
        """'pmt6' (build_8inch_pmt(nsteps=6), 1,284 triangles) or 'pmt24lc'
        (build_8inch_pmt_with_lc(nsteps=24), 5,856 triangles)."""
        z = self.z
        mesh = Mesh(z[name + '.vertices'], z[name + '.triangles'], round=False, remove_null_triangles=False)
        pick = lambda pool, idx: [None if i < 0 else pool[i] for i in idx]
        return Solid(mesh, pick(self.materials, z[name + '.material1']), pick(self.materials, z[name + '.material2']),
                     pick(self.surfaces, z[name + '.surface']), list(z[name + '.color']))

    def lion_mesh(self, subdivide=0):
        """chroma/models/lionsolid.stl.bz2 (74,358 triangles); every subdivision splits each
        triangle into four through its edge midpoints (2 -> 1,189,728 triangles)."""
        v, t = self.z['lion.vertices'].astype(np.float64), self.z['lion.triangles'].astype(np.int64)
        for _ in range(subdivide):
            v, t = subdivide_midpoints(v, t)
        return Mesh(v.astype(np.float32), t.astype(np.int32), round=False, remove_null_triangles=False)


def subdivide_midpoints(vertices, triangles):
    """1 -> 4 subdivision with shared edge midpoints."""
    nv = len(vertices)
    edges = np.sort(np.concatenate([triangles[:, [0, 1]], triangles[:, [1, 2]], triangles[:, [2, 0]]]), axis=1)
    uniq, inverse = np.unique(edges, axis=0, return_inverse=True)
    mid = 0.5 * (vertices[uniq[:, 0]] + vertices[uniq[:, 1]])
    nt = len(triangles)
    inverse = np.asarray(inverse).reshape(-1)
    m01, m12, m20 = nv + inverse[:nt], nv + inverse[nt:2 * nt], nv + inverse[2 * nt:]
    a, b, c = triangles[:, 0], triangles[:, 1], triangles[:, 2]
    tris = np.concatenate([np.stack([a, m01, m20], 1), np.stack([m01, b, m12], 1),
                           np.stack([m20, m12, c], 1), np.stack([m01, m12, m20], 1)])
    return np.concatenate([vertices, mid]), tris


def spiral_positions(radius, spacing):
    """Points `spacing` apart along a spiral wrapped on a sphere (chroma/demo/__init__.py:19-30)."""
    dl = spacing / radius
    t, a = 0.0, np.pi / dl
    while t < np.pi:
        yield np.array([sin(t) * sin(a * t), sin(t) * cos(a * t), cos(t)]) * radius
        t += dl / sqrt(1 + a ** 2 * sin(t) ** 2)


def detector(parts=None, pmt='pmt6', pmt_radius=14000.0, sphere_radius=14500.0, spiral_step=350.0, shell_nsteps=200):
    """The reference's demo detector (chroma/demo/__init__.py:32-64): water sphere with a black
    liner, inward-facing PMTs along a spiral, Gaussian time (1.5 ns) and charge (1.0 +- 0.1)
    response."""
    parts = parts or Parts()
    water = parts.material('water')
    model = parts.solid(pmt)
    geo = Detector(water)
    geo.add_solid(Solid(sphere(sphere_radius, nsteps=shell_nsteps), water, water, surface=parts.surface('black_surface'),
                        color=0xBBFFFFFF))
    facing = np.array((0.0, 1.0, 0.0))            # the PMT model looks along +y
    for position in spiral_positions(pmt_radius, spiral_step):
        inward = -normalize(position)
        turn = np.arccos(np.dot(facing, inward))
        geo.add_pmt(model, make_rotation_matrix(turn, np.cross(inward, facing)), position)
    geo.set_time_dist_gaussian(1.5, -5 * 1.5, 5 * 1.5)
    geo.set_charge_dist_gaussian(1.0, 0.1, 0.0, 1.0 + 5 * 0.1)
    geo.parts = parts
    return geo


def detector_29k(parts=None, pmt='pmt6'):
    """BASELINE config 3: 28,995 PMTs on a 23.775 m sphere; pmt='pmt24lc'
    is the heavy variant (169.8 M triangles)."""
    return detector(parts, pmt, pmt_radius=23775.0, sphere_radius=24275.0, spiral_step=350.0)


def tiny(parts=None, pmt='pmt6'):
    return detector(parts, pmt, 2000.0, 2500.0, 700.0, shell_nsteps=64)
