"""The detector BASELINE config 3 is quoted on, rebuilt from the reference's own parts
(tests/golden/ref_detector_parts.npz, generator make_ref_detector_parts.py) by this package's
placement code, pinned against what the reference's chroma.demo.detector-style builder produced
with the same parts (chroma/demo/__init__.py:19-64).  No GPU."""
import os
import numpy as np
import pytest

from chroma_lite_b200.demo import refparts

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'ref_detector_parts.npz')


@pytest.fixture(scope='module')
def parts():
    return refparts.Parts(GOLD)


def test_parts_are_the_reference_models(parts):
    pmt = parts.solid('pmt6')
    assert len(pmt.mesh.triangles) == 1284                       # build_8inch_pmt(nsteps=6), SURVEY App. D
    assert len(parts.solid('pmt24lc').mesh.triangles) == 5856    # build_8inch_pmt_with_lc(nsteps=24)
    names = sorted(m.name for m in pmt.unique_materials)
    assert names == ['glass', 'vacuum', 'water']
    assert sorted(s.name for s in pmt.unique_surfaces if s is not None) == ['r7081hqe_photocathode', 'shiny_surface']
    qe = parts.surface('r7081hqe_photocathode').detect
    assert qe.shape[1] == 2 and 0.2 < qe[:, 1].max() < 0.4      # R7081HQE quantum efficiency peaks near 0.3
    water = parts.material('water')
    assert abs(np.interp(400.0, water.refractive_index[:, 0], water.refractive_index[:, 1]) - 1.34) < 0.02
    assert len(parts.lion_mesh().triangles) == 74358
    assert sum(1 for _ in refparts.spiral_positions(23775.0, 350.0)) == int(parts.z['check.npmt_29k']) == 28995


def test_placement_matches_the_reference_builder(parts):
    z = parts.z
    det = refparts.tiny(parts)
    det.flatten()
    assert len(det.mesh.vertices) == int(z['check.tiny_nvertices'])
    assert len(det.mesh.triangles) == int(z['check.tiny_ntriangles'])
    assert det.num_channels() == int(z['check.tiny_nchannels'])
    assert np.array_equal(det.mesh.vertices[::997], z['check.tiny_rows'])                 # bit for bit
    assert np.array_equal(det.mesh.triangles[::997].astype(np.uint32), z['check.tiny_triangle_rows'])
    v = det.mesh.vertices.astype(np.float64)
    assert np.allclose(v.sum(axis=0), z['check.tiny_vertex_sum'], rtol=0, atol=1e-6)
    assert np.allclose(np.abs(v).sum(axis=0), z['check.tiny_vertex_abs_sum'], rtol=0, atol=1e-6)
    # per-triangle media: the same number of triangles per material / surface NAME (index order may differ)
    by_name = lambda objs, idx: {getattr(o, 'name', None): int((idx == i).sum()) for i, o in enumerate(objs)}
    m1 = by_name(det.unique_materials, det.material1_index)
    assert [m1[n] for n in ('vacuum', 'glass', 'water')] == list(z['check.tiny_material1'])
    surf = by_name(det.unique_surfaces, det.surface_index)
    # the fixture holds bincount(surface_index + 1) of the reference's flatten, whose surface list is
    # [None, shiny, black, photocathode] with None stored as -1: slots [None, (unused 0), shiny, black, photocathode]
    c = z['check.tiny_surface']
    assert len(c) == 5 and c[1] == 0
    ref = {None: c[0], 'shiny_surface': c[2], 'black_surface': c[3], 'r7081hqe_photocathode': c[4]}
    assert int((det.surface_index == -1).sum()) == ref[None]
    for n in ('shiny_surface', 'black_surface', 'r7081hqe_photocathode'):
        assert surf[n] == ref[n]
    assert np.allclose(det.time_cdf[0], z['check.time_cdf_x']) and len(det.time_cdf[1]) == len(det.time_cdf[0])


def test_subdivision_keeps_the_surface(parts):
    m0 = parts.lion_mesh()
    m1 = parts.lion_mesh(subdivide=1)
    assert len(m1.triangles) == 4 * len(m0.triangles)
    area = lambda m: 0.5 * np.linalg.norm(np.cross(m.vertices[m.triangles[:, 1]].astype(np.float64) - m.vertices[m.triangles[:, 0]],
                                                   m.vertices[m.triangles[:, 2]].astype(np.float64) - m.vertices[m.triangles[:, 0]]), axis=1).sum()
    assert abs(area(m1) / area(m0) - 1.0) < 1e-5
