"""Demo geometries used by the tests and the benchmark (role of chroma/demo):
a PMT built as a surface of revolution, a spiral-tiled spherical
water-Cherenkov detector (chroma/demo/__init__.py:19-64) and the acrylic-sphere
scene of BASELINE config 1."""
from math import sin, cos, sqrt

import numpy as np

from ..geometry import Solid, Geometry
from ..detector import Detector
from ..make import rotate_extrude, sphere
from ..sample import make_rotation_matrix, normalize
from . import optics


def pmt_profile(npoints=55, radius=101.0, neck_radius=42.0, depth=75.0, neck_length=110.0):
    """Outline of an 8-inch-class PMT facing +y: a cylindrical neck closed at the
    base, flaring into an oblate bulb; returned bottom (x=0) to top (x=0)."""
    nbulb = npoints - 8
    yb = -depth - neck_length
    neck = [(0.0, yb), (neck_radius * 0.6, yb), (neck_radius, yb + 4.0),
            (neck_radius, yb + 0.35 * neck_length), (neck_radius, yb + 0.7 * neck_length),
            (neck_radius * 1.05, -depth - 0.1 * neck_length)]
    # bulb: ellipse from polar angle ~115 deg (joins the neck) to the pole
    t = np.linspace(np.radians(128.0), 0.0, nbulb + 2)
    bulb = [(radius * np.sin(a), depth * np.cos(a)) for a in t]
    prof = np.array(neck + bulb, dtype=np.float64)
    prof[0, 0] = 0.0
    prof[-1, 0] = 0.0
    return prof


def _offset_profile(profile, d):
    """Shrink a closed-at-the-axis profile by d along the local normal."""
    p = np.asarray(profile, dtype=np.float64)
    tang = np.gradient(p, axis=0)
    tang /= np.maximum(np.linalg.norm(tang, axis=1), 1e-12)[:, None]
    normal = np.stack([tang[:, 1], -tang[:, 0]], axis=1)     # outward for a CCW profile
    q = p - d * normal
    q[:, 0] = np.maximum(q[:, 0], 0.0)
    q[0, 0] = 0.0
    q[-1, 0] = 0.0
    q[0, 1] = p[0, 1] + d
    q[-1, 1] = p[-1, 1] - d
    return q


def build_pmt(nsteps=6, npoints=55, glass_thickness=3.0, outer_material=None):
    """Glass envelope in `outer_material` with a vacuum cavity whose upper half
    carries the photocathode and lower half a mirror (structure of
    chroma/pmt.py:40-73).  nsteps=6, npoints=55 -> 1,284 triangles."""
    outer_material = optics.water if outer_material is None else outer_material
    prof = pmt_profile(npoints)
    inner = _offset_profile(prof, glass_thickness)
    outer_mesh = rotate_extrude(prof[:, 0], prof[:, 1], nsteps)
    inner_mesh = rotate_extrude(inner[:, 0], inner[:, 1], nsteps)
    outer = Solid(outer_mesh, optics.glass, outer_material, color=0xeeffffff)
    front = np.mean(inner_mesh.assemble(), axis=1)[:, 1] > 0
    surf = np.where(front, optics.photocathode, optics.shiny_surface)
    inner_solid = Solid(inner_mesh, optics.vacuum, optics.glass, surface=list(surf),
                        color=list(np.where(front, 0xff00, 0xff0000)))
    return outer + inner_solid


def spherical_spiral(radius, spacing):
    """Points roughly `spacing` apart along a spiral wrapped on a sphere."""
    dl = spacing / radius
    t = 0.0
    a = np.pi / dl
    while t < np.pi:
        yield np.array([sin(t) * sin(a * t), sin(t) * cos(a * t), cos(t)]) * radius
        t += dl / sqrt(1 + a ** 2 * sin(t) ** 2)


def detector(pmt_radius=14000.0, sphere_radius=14500.0, spiral_step=350.0, pmt_nsteps=6,
             shell_nsteps=200, max_pmts=None):
    """Water sphere with a black liner tiled with inward-facing PMTs."""
    pmt = build_pmt(nsteps=pmt_nsteps)
    geo = Detector(optics.water)
    geo.add_solid(Solid(sphere(sphere_radius, nsteps=shell_nsteps), optics.water, optics.water,
                        surface=optics.black_surface, color=0xBBFFFFFF))
    y_axis = np.array((0.0, 1.0, 0.0))
    for k, position in enumerate(spherical_spiral(pmt_radius, spiral_step)):
        if max_pmts is not None and k >= max_pmts:
            break
        direction = -normalize(position)
        axis = np.cross(direction, y_axis)
        if np.linalg.norm(axis) < 1e-9:
            axis = np.array((1.0, 0.0, 0.0))
        angle = np.arccos(np.clip(np.dot(y_axis, direction), -1.0, 1.0))
        # rotate the PMT's +y axis onto `direction`
        geo.add_pmt(pmt, make_rotation_matrix(-angle, axis), position)
    time_rms, charge_mean, charge_rms = 1.5, 1.0, 0.1
    geo.set_time_dist_gaussian(time_rms, -5 * time_rms, 5 * time_rms)
    geo.set_charge_dist_gaussian(charge_mean, charge_rms, 0.0, charge_mean + 5 * charge_rms)
    return geo


def detector_29k(pmt_nsteps=6):
    """BASELINE config 3: 28,995 PMTs on a 23.775 m sphere."""
    return detector(pmt_radius=23775.0, sphere_radius=24275.0, spiral_step=350.0, pmt_nsteps=pmt_nsteps)


def tiny(pmt_nsteps=6):
    return detector(2000.0, 2500.0, 700.0, pmt_nsteps=pmt_nsteps, shell_nsteps=64)


def acrylic_sphere_scene(nsteps=64):
    """BASELINE config 1: acrylic sphere (R = 1 m) in water inside a black shell (R = 5 m)."""
    geo = Geometry(optics.water)
    geo.add_solid(Solid(sphere(1000.0, nsteps), optics.acrylic, optics.water))
    geo.add_solid(Solid(sphere(5000.0, nsteps), optics.water, optics.water, surface=optics.black_surface))
    return geo
