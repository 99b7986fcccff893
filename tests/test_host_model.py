"""Host-side data model: flatten() against the reference's own output (golden),
BVH packing known answers (reference test/test_bvh.py), table pool, helpers."""
import os
import numpy as np
import pytest

from chroma_lite_b200 import event, make
from chroma_lite_b200.bvh import BVH, WorldCoords, unpack_nodes, uint4
from chroma_lite_b200.detector import Detector
from chroma_lite_b200.geometry import Mesh, Solid, Material, Surface, Geometry
from chroma_lite_b200.gpu.tools import chunk_iterator, to_float3
from chroma_lite_b200.gpu.geometry import make_desc, material_codes, interp_property
import scenes

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def test_flatten_matches_reference_golden():
    z = np.load(os.path.join(GOLD, 'flatten.npz'))
    m_a, m_b, s_x = Material('a'), Material('b'), Surface('x')
    det = Detector(m_a)
    shell = Mesh(z['shell_vertices'], z['shell_triangles'], round=False, remove_null_triangles=False)
    small = Mesh(z['small_vertices'], z['small_triangles'], round=False, remove_null_triangles=False)
    det.add_solid(Solid(shell, m_a, m_a, surface=s_x))
    rot = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]], dtype=np.float32)
    for k in range(3):
        det.add_pmt(Solid(small, m_b, m_a), rotation=rot if k % 2 else None, displacement=(100.0 * k, 5.0, -20.0 * k))
    det.flatten()
    assert np.array_equal(det.mesh.vertices, z['vertices'])
    assert np.array_equal(det.mesh.triangles, z['triangles'])
    assert np.array_equal(det.solid_id, z['solid_id'])
    assert np.array_equal(det.colors, z['colors'])
    assert np.array_equal(np.array([det.unique_materials[i] is m_b for i in det.material1_index]), z['material1_is_b'])
    assert np.array_equal(np.array([det.unique_materials[i] is m_b for i in det.material2_index]), z['material2_is_b'])
    assert np.array_equal(det.surface_index >= 0, z['surface_is_x'])
    assert np.array_equal(det.solid_id_to_channel_index, z['solid_id_to_channel_index'])


def test_sphere_builder_same_shape_as_reference():
    z = np.load(os.path.join(GOLD, 'sphere_mesh.npz'))
    m = make.sphere(1000.0, 16)
    assert len(m.triangles) == len(z['triangles']) and len(m.vertices) == len(z['vertices'])
    r = np.linalg.norm(m.vertices, axis=1)
    assert np.allclose(r, 1000.0, rtol=1e-5)
    # outward winding
    tri = m.assemble()
    n = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 1])
    assert (np.einsum('ij,ij->i', n, tri.mean(axis=1)) > 0).all()


def _ref_test_bvh():
    # the 3-layer binary tree of the reference's test/test_bvh.py:54-92
    nodes = np.empty(7, dtype=uint4)
    nodes['x'][3:7] = [0x00010000, 0x00020001, 0x00010000, 0x00010000]
    nodes['y'][3:7] = [0x00010000, 0x00010000, 0x00020001, 0x00010000]
    nodes['z'][3:7] = [0x00010000, 0x00010000, 0x00010000, 0x00020001]
    nodes['w'][3:7] = 0x80000000
    nodes['x'][1:3] = [0x00020000, 0x00010000]
    nodes['y'][1:3] = [0x00010000, 0x00020000]
    nodes['z'][1:3] = [0x00010000, 0x00020000]
    nodes['w'][1:3] = [0x00000003, 0x00000005]
    nodes['x'][0:1], nodes['y'][0:1], nodes['z'][0:1], nodes['w'][0:1] = 0x00020000, 0x00020000, 0x00020000, 1
    return BVH(WorldCoords(np.array([-1.0, -1.0, -1.0]), 0.1), nodes, [0, 1, 3])


def test_unpack_nodes_known_answers():
    bvh = _ref_test_bvh()
    u = unpack_nodes(bvh.get_layer(2).nodes)
    assert list(u['xlo']) == [0, 1, 0, 0] and list(u['xhi']) == [1, 2, 1, 1]
    assert list(u['ylo']) == [0, 0, 1, 0] and list(u['yhi']) == [1, 1, 2, 1]
    assert list(u['zlo']) == [0, 0, 0, 1] and list(u['zhi']) == [1, 1, 1, 2]
    u = unpack_nodes(bvh.get_layer(1).nodes)
    assert list(u['xhi']) == [2, 1] and list(u['yhi']) == [1, 2] and list(u['child']) == [3, 5]
    u = unpack_nodes(bvh.get_layer(0).nodes)
    assert list(u['xhi']) == [2] and list(u['child']) == [1]
    assert len(bvh) == 7 and bvh.layer_count() == 3
    assert [len(bvh.get_layer(i)) for i in range(3)] == [1, 2, 4]
    # layer areas in world units (the reference's test_bvh.py:148-156)
    assert np.isclose(bvh.get_layer(2).area(), 4 * 6 * 0.1 ** 2)
    assert np.isclose(bvh.get_layer(1).area(), (4 * 2 + 2 + 4 * 2 + 2 * 4) * 0.1 ** 2)
    assert np.isclose(bvh.get_layer(0).area(), 6 * 2 * 2 * 0.1 ** 2)
    lo, hi = bvh.get_layer(0).get_bounds()
    assert np.allclose(lo, [[-1, -1, -1]]) and np.allclose(hi, [[-0.8, -0.8, -0.8]])


def test_world_coords_known_answers():
    wc = WorldCoords(world_origin=[-1, -1, -1], world_scale=0.1)
    w = [[-1.0, -0.9, 9.0], [1.0, 3.0, 5.0], [20.0, 30.0, 40.0]]
    assert np.array_equal(wc.world_to_fixed(w), [[0, 1, 100], [20, 40, 60], [210, 310, 410]])
    assert np.allclose(wc.fixed_to_world([[0, 1, 100]]), [[-1.0, -0.9, 9.0]], atol=1e-6)
    from chroma_lite_b200.bvh import OutOfRangeError
    with pytest.raises(OutOfRangeError):
        wc.world_to_fixed([-2.0, 0.0, 0.0])
    with pytest.raises(OutOfRangeError):
        wc.world_to_fixed([0.0, 1e9, 0.0])


def test_chunk_iterator_matches_reference_semantics():
    assert list(chunk_iterator(300, 32, 2)) == [(0, 64, 2), (64, 64, 2), (128, 64, 2), (192, 64, 2), (256, 44, 2)]
    assert list(chunk_iterator(0)) == []
    assert list(chunk_iterator(64, 64, 1024)) == [(0, 64, 1)]
    assert sum(n for _, n, _ in chunk_iterator(1234567, 512, 1024)) == 1234567


def test_material_codes_sign_extension():
    g = type('G', (), {})()
    g.material1_index = np.array([0, 1, 5])
    g.material2_index = np.array([2, 0, 127])
    g.surface_index = np.array([-1, 3, 0])
    c = material_codes(g)
    assert list(c >> 24) == [0, 1, 5] and list((c >> 16) & 0xff) == [2, 0, 127]
    assert list((c >> 8) & 0xff) == [0xff, 3, 0]        # -1 -> 0xff, decoded back by sign extension


def test_table_pool_layout_and_values():
    geo = scenes.scintillator_scene(8)
    desc, keep = make_desc(geo)
    pool, mats, surfs = keep['pool'], keep['mats'], keep['surfs']
    W, T = desc.wavelength_n, desc.time_n
    assert W == 188 and T == 20000 and desc.wavelength_start == 60.0 and desc.wavelength_step == 5.0
    wl = np.arange(60, 1000, 5).astype(np.float32)
    for i, m in enumerate(geo.unique_materials):
        cm = mats[i]
        assert np.array_equal(pool[cm.refractive_index:cm.refractive_index + W], interp_property(wl, m.refractive_index))
        assert cm.num_comp == len(m.comp_reemission_prob)
        if cm.num_comp:
            assert cm.comp_reemission_time_cdf >= cm.comp_absorption_length   # time CDFs live at the back
            tc = pool[cm.comp_reemission_time_cdf:cm.comp_reemission_time_cdf + cm.num_comp * T]
            assert np.isclose(tc[T - 1], 1.0) and np.isclose(tc[2 * T - 1], 1.0)
    models = sorted(s.model for s in surfs[:desc.nsurfaces])
    assert models == sorted([-1] + [s.model for s in geo.unique_surfaces if s is not None])
    for i, s in enumerate(geo.unique_surfaces):
        if s is not None and s.dichroic_props is not None:
            cs = surfs[i]
            assert cs.dichroic_nangles == 4
            assert np.allclose(pool[cs.dichroic_angles:cs.dichroic_angles + 4], s.dichroic_props.angles)
    # every time CDF lies after every wavelength table (the engine stages the front in shared memory)
    first_time = min(m.comp_reemission_time_cdf for m in mats[:desc.nmaterials] if m.num_comp)
    assert first_time + sum(m.num_comp for m in mats[:desc.nmaterials]) * T == len(pool)


def test_photons_event_contract():
    ph = scenes.point_source(10)
    assert ph.pos.dtype == np.float32 and ph.flags.dtype == np.uint32 and ph.last_hit_triangles.dtype == np.int32
    assert (ph.last_hit_triangles == -1).all() and (ph.weights == 1).all()
    both = ph + ph
    assert len(both) == 20 and len(event.Photons.join([ph, ph, ph])) == 30
    assert len(ph[ph.wavelengths > 0]) == 10 and len(ph[2:5]) == 3
    assert event.NAN_ABORT == 1 << 31 and event.NAN_ABORT_KERNEL == 1 << 15
    f3 = to_float3(ph.pos)
    assert f3.shape == (10,) and f3['z'][3] == ph.pos[3, 2]
    ch = event.Channels(np.array([True, False]), np.array([1.0, 1e9]), np.array([1.0, 0.0]), np.array([4, 0]))
    ids, t, q = ch.hit_channels()
    assert list(ids) == [0]


def test_native_vertex_dedup_equals_numpy(monkeypatch):
    """cb_unique_vertices (multi-threaded host sort, csrc/hostmesh.cu) gives np.unique's rows and
    inverse; Mesh uses it for large meshes (chroma/geometry.py:59-69)."""
    from chroma_lite_b200 import geometry as G
    rng = np.random.default_rng(3)
    for n in (1, 2, 1000, 300000):
        v = (rng.integers(-40, 40, (n, 3)) / 8.0).astype(np.float32)
        v[::5] *= np.float32(-0.0) if n > 2 else 1          # signed zeros compare equal, as in NumPy
        uniq, inverse = G._native_unique_vertices(v)
        assert uniq is not None
        ref_u, ref_i = np.unique(v, axis=0, return_inverse=True)
        assert np.array_equal(uniq, ref_u) and np.array_equal(inverse, np.asarray(ref_i).reshape(-1))
        assert np.array_equal(uniq[inverse], v)
    # the same mesh through both paths
    from chroma_lite_b200.make import sphere
    m = sphere(10.0, 96)
    tiled = np.concatenate([m.vertices + np.float32(k % 3) for k in range(12)])
    tris = np.concatenate([m.triangles + k * len(m.vertices) for k in range(12)])
    a = G.Mesh(tiled, tris, remove_duplicate_vertices=True, remove_null_triangles=False)
    monkeypatch.setattr(G, 'NATIVE_UNIQUE_MIN', 1)
    b = G.Mesh(tiled, tris, remove_duplicate_vertices=True, remove_null_triangles=False)
    assert len(a.vertices) < len(tiled)
    assert np.array_equal(a.vertices, b.vertices) and np.array_equal(a.triangles, b.triangles)


def test_response_cdfs_are_uploaded_with_equal_lengths():
    """gpu/detector.py: the reference's _pdf_to_cdf adds 0.0 to the cumulative sum instead of prepending it
    (chroma/detector.py:104-107), so its y array is one short of x; the upload completes such a pair and
    rejects anything else instead of reading past the end."""
    import pytest
    from chroma_lite_b200.gpu.detector import cdf_arrays
    from chroma_lite_b200.detector import Detector
    d = Detector(None)
    d.set_time_dist_gaussian(1.2, -6.0, 6.0)
    x, y = cdf_arrays(d.time_cdf)
    assert len(x) == len(y) == 51 and y[0] == 0.0 and y[-1] == 1.0 and (np.diff(y) >= 0).all()
    # the pair the reference's Detector produces for the same call
    edges = np.linspace(-6.0, 6.0, 51)
    contents = np.exp(-0.5 * (edges[1:] / 1.2) ** 2)
    ref_y = np.array([0.0] + contents.cumsum())
    ref_y /= ref_y[-1]
    assert len(ref_y) == 50
    x2, y2 = cdf_arrays((edges, ref_y))
    assert np.array_equal(x2, x) and np.allclose(y2, y, atol=1e-7)
    with pytest.raises(ValueError):
        cdf_arrays((edges, ref_y[:-3]))
    with pytest.raises(ValueError):
        cdf_arrays((edges[:1], ref_y[:1]))


def test_transform_builders_and_flashlight_match_the_reference():
    """tests/golden/host_helpers.npz (made from the reference's chroma.transform / chroma.make / chroma.sample by
    tests/golden/make_golden.py): same rotation sense and values, same meshes up to vertex numbering (triangle and
    vertex counts, area, signed volume, so orientation too), same flashlight sample under the same NumPy seed."""
    import os
    import sys
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    g = np.load(os.path.join(here, 'host_helpers.npz'))
    from chroma_lite_b200 import transform, make, sample
    axis, phi, x = np.array([0.3, -0.5, 0.8]), 0.7, g['points']
    assert np.allclose(transform.make_rotation_matrix(phi, axis), g['matrix'], atol=1e-14)
    assert np.allclose(transform.rotate(x, phi, axis), g['rotate'], atol=1e-13)
    assert np.allclose(transform.rotate(x, np.linspace(0, 1, 5), axis), g['rotate_many'], atol=1e-13)
    assert np.allclose(transform.rotate_matrix(x, phi, axis), g['rotate_matrix'], atol=1e-13)
    assert np.allclose(np.inner(x, transform.make_rotation_matrix(phi, axis)), g['rotate'], atol=1e-13)
    assert np.allclose(transform.get_perp(axis), g['get_perp'])
    pairs = (([1, 0, 0], [0, 1, 0]), ([1, 0, 0], [1, 0, 0]), ([0, 1, 0], [0, 1, 0]), ([0, 0, 1], [0, 0, -1]),
             ([1, 2, 3], [-2, 0.5, 1]))
    for (a, b), want in zip(pairs, g['gen_rot']):
        assert np.allclose(transform.gen_rot(a, b), want, atol=1e-13)
    # the right-handed helper of this package is the transpose (what demo.detector compensates with -angle)
    assert np.allclose(sample.make_rotation_matrix(phi, axis), g['matrix'].T, atol=1e-14)

    ang = np.linspace(0, 2 * np.pi, 6, endpoint=False)
    hexagon = (np.cos(ang), np.sin(ang))
    cases = (('linear_extrude', hexagon + (2.0,), {}),
             ('linear_extrude', hexagon + (2.0,), dict(x2=0.5 * hexagon[0], y2=0.5 * hexagon[1], center=(1, 2, 3))),
             ('linear_extrude', hexagon + (2.0,), dict(endcaps=False)),
             ('cylinder_along_z', (10.0, 30.0, 20), {}), ('segmented_cylinder', (10.0, 30.0, 16, 40), {}),
             ('torus', (2.0, 10.0, 16, 12), {}), ('convex_polygon', hexagon, {}), ('cylinder', (5.0, 8.0, 3.0, 12), {}))
    for (name, args, kw), want in zip(cases, g['builders']):
        m = getattr(make, name)(*args, **kw)
        v = np.asarray(m.vertices, dtype=np.float64)[np.asarray(m.triangles)]
        cr = np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0])
        got = [len(m.triangles), len(m.vertices), 0.5 * np.linalg.norm(cr, axis=1).sum(),
               np.einsum('ij,ij->i', v[:, 0], cr).sum() / 6.0]
        assert got[0] == want[0] and got[1] == want[1], name
        assert np.allclose(got[2:], want[2:], rtol=1e-9, atol=1e-9), name
    with pytest.raises(Exception):
        make.linear_extrude([0, 1], [0, 1, 2], 1.0)
    # a closed prism: hexagon area * height
    hexprism = make.linear_extrude(hexagon[0], hexagon[1], 2.0)
    assert np.isclose(g['builders'][0][3], 2.0 * 1.5 * np.sqrt(3.0), rtol=1e-6) and len(hexprism.md5()) == 32

    np.random.seed(5)
    assert np.allclose(sample.flashlight(0.3, (1, 2, 3), 1000), g['flashlight'], atol=1e-12)
    np.random.seed(5)
    assert np.allclose(sample.flashlight(), g['flashlight_one'], atol=1e-12)
    d = sample.flashlight(0.2, (0, 1, 0), 500, rng=np.random.default_rng(1))
    assert np.allclose(np.linalg.norm(d, axis=1), 1.0) and (d[:, 1] >= np.cos(0.2) - 1e-12).all()
