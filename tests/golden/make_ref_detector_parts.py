"""Generate tests/golden/ref_detector_parts.npz: the building blocks of the reference's own
demo detector, produced by the reference's own pure-Python modules.

Run HERE (container with /root/reference):  python tests/golden/make_ref_detector_parts.py

What is stored (BASELINE configs 2 and 3; VERDICT r01 "What's missing" #2):
  pmt6.*      chroma.demo.pmt.build_8inch_pmt(nsteps=6)          1,280 triangles
  pmt24lc.*   chroma.demo.pmt.build_8inch_pmt_with_lc(nsteps=24) 5,856 triangles (heavy variant)
              per solid: vertices f32, triangles, per-triangle material1 / material2 / surface
              (indices into the name lists below, -1 = no surface), colours
  mat.<name>.<property>, surf.<name>.<property>
              the (wavelength, value) tables of chroma/demo/optics.py: water (WCSim), glass (SNO),
              vacuum, r7081hqe_photocathode, shiny_surface, black_surface
  lion.*      chroma/models/lionsolid.stl.bz2 through chroma.stl.mesh_from_stl (config 2's mesh)
  check.*     pins for the host-side placement: the reference's chroma.demo.detector(2000, 2500, 700)
              built with the nsteps=6 PMT and flattened (vertex / triangle counts, float64 sums of
              the vertex array and of a few strided rows) and its spiral's PMT count for the 29k
              detector, so that this package's builder can be checked against the reference's without
              the reference being importable on the GPU box.
Nothing of the reference's GPU host code is imported (it needs PyCUDA).
"""
import os
import sys
import numpy as np

REF = os.environ.get('CHROMA_REFERENCE', '/root/reference')
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))

import chroma.demo.optics as roptics               # noqa: E402
import chroma.demo.pmt as rpmt                     # noqa: E402
import chroma.demo as rdemo                        # noqa: E402
from chroma.geometry import Solid                  # noqa: E402
from chroma.detector import Detector               # noqa: E402
from chroma.make import sphere                     # noqa: E402
from chroma.stl import mesh_from_stl               # noqa: E402
from chroma.transform import make_rotation_matrix, normalize   # noqa: E402

MATERIALS = ['water', 'glass', 'vacuum']
SURFACES = ['r7081hqe_photocathode', 'shiny_surface', 'black_surface']
MAT_PROPS = ['refractive_index', 'absorption_length', 'scattering_length']
SURF_PROPS = ['detect', 'absorb', 'reemit', 'reflect_diffuse', 'reflect_specular', 'eta', 'k', 'reemission_cdf']


def solid_arrays(prefix, solid, out):
    mats = [getattr(roptics, n) for n in MATERIALS]
    surfs = [getattr(roptics, n) for n in SURFACES]

    def index(objs, pool):
        return np.array([-1 if o is None else [id(x) for x in pool].index(id(o)) for o in objs], dtype=np.int32)
    out[prefix + '.vertices'] = np.asarray(solid.mesh.vertices, dtype=np.float32)
    out[prefix + '.triangles'] = np.asarray(solid.mesh.triangles, dtype=np.uint32)
    out[prefix + '.material1'] = index(solid.material1, mats)
    out[prefix + '.material2'] = index(solid.material2, mats)
    out[prefix + '.surface'] = index(solid.surface, surfs)
    out[prefix + '.color'] = np.asarray(solid.color, dtype=np.uint32)


def reference_detector(pmt, pmt_radius, sphere_radius, spiral_step, shell_nsteps=200):
    """chroma/demo/__init__.py:32-64 with the PMT model as a parameter (the reference hard-codes
    build_8inch_pmt_with_lc(); BASELINE config 3 uses the nsteps=6 PMT)."""
    geo = Detector(roptics.water)
    geo.add_solid(Solid(sphere(sphere_radius, nsteps=shell_nsteps), roptics.water, roptics.water,
                        surface=roptics.black_surface, color=0xBBFFFFFF))
    for position in rdemo.spherical_spiral(pmt_radius, spiral_step):
        direction = -normalize(position)
        y_axis = np.array((0.0, 1.0, 0.0))
        axis = np.cross(direction, y_axis)
        angle = np.arccos(np.dot(y_axis, direction))
        geo.add_pmt(pmt, make_rotation_matrix(angle, axis), position)
    geo.set_time_dist_gaussian(1.5, -7.5, 7.5)
    geo.set_charge_dist_gaussian(1.0, 0.1, 0.0, 1.5)
    return geo


def main():
    out = {'materials': np.array(MATERIALS), 'surfaces': np.array(SURFACES)}
    for n in MATERIALS:
        m = getattr(roptics, n)
        for p in MAT_PROPS:
            out['mat.%s.%s' % (n, p)] = np.asarray(getattr(m, p), dtype=np.float32)
        out['mat.%s.density' % n] = np.float64(getattr(m, 'density', 0.0))
    for n in SURFACES:
        s = getattr(roptics, n)
        for p in SURF_PROPS:
            out['surf.%s.%s' % (n, p)] = np.asarray(getattr(s, p), dtype=np.float32)
        out['surf.%s.model' % n] = np.int32(s.model)
        out['surf.%s.thickness' % n] = np.float32(s.thickness)
        out['surf.%s.transmissive' % n] = np.int32(s.transmissive)
    pmt6 = rpmt.build_8inch_pmt(nsteps=6)
    solid_arrays('pmt6', pmt6, out)
    pmt24 = rpmt.build_8inch_pmt_with_lc(nsteps=24)
    solid_arrays('pmt24lc', pmt24, out)
    lion = mesh_from_stl(os.path.join(REF, 'chroma', 'models', 'lionsolid.stl.bz2'))
    out['lion.vertices'] = np.asarray(lion.vertices, dtype=np.float32)
    out['lion.triangles'] = np.asarray(lion.triangles, dtype=np.uint32)

    # pins for the placement code
    npmt = sum(1 for _ in rdemo.spherical_spiral(23775.0, 350.0))
    out['check.npmt_29k'] = np.int64(npmt)
    det = reference_detector(pmt6, 2000.0, 2500.0, 700.0, shell_nsteps=64)
    det.flatten()
    v = det.mesh.vertices.astype(np.float64)
    out['check.tiny_nvertices'] = np.int64(len(v))
    out['check.tiny_ntriangles'] = np.int64(len(det.mesh.triangles))
    out['check.tiny_vertex_sum'] = v.sum(axis=0)
    out['check.tiny_vertex_abs_sum'] = np.abs(v).sum(axis=0)
    out['check.tiny_rows'] = det.mesh.vertices[::997].astype(np.float32)
    out['check.tiny_triangle_rows'] = det.mesh.triangles[::997].astype(np.uint32)
    out['check.tiny_material1'] = np.bincount(det.material1_index, minlength=3).astype(np.int64)
    out['check.tiny_surface'] = np.bincount(det.surface_index + 1, minlength=4).astype(np.int64)
    out['check.tiny_nchannels'] = np.int64(det.num_channels())
    out['check.time_cdf_x'] = np.asarray(det.time_cdf[0], dtype=np.float64)
    out['check.time_cdf_y'] = np.asarray(det.time_cdf[1], dtype=np.float64)
    path = os.path.join(HERE, 'ref_detector_parts.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes;', 'pmt6', len(pmt6.mesh.triangles), 'pmt24lc',
          len(pmt24.mesh.triangles), 'lion', len(lion.triangles), 'npmt', npmt,
          'materials of the tiny detector', [m.name for m in det.unique_materials],
          'surfaces', [s.name if s is not None else None for s in det.unique_surfaces])


if __name__ == '__main__':
    main()
