"""Binary BVH / flattened-mesh cache (role of chroma/cache.py), no GPU."""
import numpy as np
import pytest

from chroma_lite_b200 import cache
from chroma_lite_b200.bvh import BVH, WorldCoords, uint4
import scenes


def fake_bvh(n=1000):
    nodes = np.zeros(n, dtype=uint4)
    raw = nodes.view(np.uint32).reshape(-1, 4)
    raw[:] = np.random.default_rng(1).integers(0, 2 ** 32, (n, 4), dtype=np.uint64).astype(np.uint32)
    return BVH(WorldCoords(np.array([-1.0, 2.0, 3.5], np.float32), np.float32(0.73)), nodes, [0, 1, 9, 200])


def test_bvh_round_trip_and_lookup_rule(tmp_path):
    c = cache.Cache(str(tmp_path / 'c'))
    bvh = fake_bvh()
    h = 'a' * 32
    assert not c.exist_bvh(h) and c.list_bvh(h) == []
    with pytest.raises(cache.BVHNotFoundError):
        c.load_bvh(h)
    c.save_bvh(bvh, h)
    c.save_bvh(bvh, h, name='degree4')
    assert c.exist_bvh(h) and c.list_bvh(h) == ['default', 'degree4']
    for mmap in (True, False):
        back = c.load_bvh(h, mmap=mmap)
        assert np.array_equal(np.asarray(back.nodes).view(np.uint32), bvh.nodes.view(np.uint32))
        assert back.layer_offsets == bvh.layer_offsets and back.layer_count() == 4
        assert np.array_equal(back.world_coords.world_origin, bvh.world_coords.world_origin)
        assert back.world_coords.world_scale == bvh.world_coords.world_scale
    c.save_bvh(fake_bvh(500), h)                 # overwrite
    assert len(c.load_bvh(h)) == 500
    c.remove_bvh(h, 'degree4')
    assert c.list_bvh(h) == ['default']


def test_mesh_hash_is_the_reference_md5():
    import hashlib
    geo = scenes.sphere_scene.__globals__['Geometry'](scenes.optics.water)
    from chroma_lite_b200.make import sphere
    m = sphere(10.0, 12)
    ref = hashlib.md5(m.vertices)
    ref.update(m.triangles)
    assert cache.mesh_hash(m) == ref.hexdigest()


def test_flattened_geometry_round_trip(tmp_path):
    from chroma_lite_b200 import demo
    det = demo.tiny()
    c = cache.Cache(str(tmp_path / 'c'))
    c.save_geometry('tiny', det)
    assert c.list_geometry() == ['tiny'] and c.get_geometry_hash('tiny') == cache.mesh_hash(det.mesh)
    back = c.load_geometry('tiny')
    assert np.array_equal(back['vertices'], det.mesh.vertices) and np.array_equal(back['triangles'], det.mesh.triangles)
    for f in ('colors', 'solid_id', 'material1_index', 'material2_index', 'surface_index'):
        assert np.array_equal(back[f], getattr(det, f))
    c.remove_geometry('tiny')
    with pytest.raises(cache.GeometryNotFoundError):
        c.load_geometry('tiny')
