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


def test_cache_directory_rules(tmp_path):
    """The reference's test_cache.py:31-70: the directory is created when missing, reused when
    present, and a plain file in its place is an error."""
    import os
    d = tmp_path / 'newdir'
    cache.Cache(str(d))
    assert os.path.isdir(d / 'geo') and os.path.isdir(d / 'bvh')
    (d / 'geo' / 'keep').mkdir()
    cache.Cache(str(d))
    assert os.path.isdir(d / 'geo' / 'keep')
    f = tmp_path / 'afile'
    f.write_text('x')
    with pytest.raises(IOError):
        cache.Cache(str(f))


def test_default_geometry_and_replacement(tmp_path):
    """test_cache.py:116-178: replacing an entry, hashes of missing entries, the '.default' link and
    a non-link in its place."""
    import os
    from chroma_lite_b200 import demo
    c = cache.Cache(str(tmp_path / 'c'))
    with pytest.raises(cache.GeometryNotFoundError):
        c.get_geometry_hash('nothing')
    with pytest.raises(cache.GeometryNotFoundError):
        c.load_default_geometry()
    with pytest.raises(cache.GeometryNotFoundError):
        c.set_default_geometry('nothing')
    a, b = demo.tiny(), demo.acrylic_sphere_scene(8)
    c.save_geometry('one', a)
    c.save_geometry('two', b)
    c.set_default_geometry('one')
    assert c.load_default_geometry()['mesh_hash'] == cache.mesh_hash(a.mesh)
    c.set_default_geometry('two')                 # the link is replaced
    assert c.load_default_geometry()['mesh_hash'] == cache.mesh_hash(b.mesh)
    c.save_geometry('two', a)                     # replace the entry behind the link
    assert c.get_geometry_hash('two') == cache.mesh_hash(a.mesh)
    assert len(c.load_default_geometry()['triangles']) == len(a.mesh.triangles)
    c.remove_geometry('does-not-exist')           # no action, no error
    os.remove(c.get_geometry_filename('.default'))
    os.mkdir(c.get_geometry_filename('.default'))
    with pytest.raises(IOError):
        c.set_default_geometry('one')


def test_loader_uses_and_fills_the_bvh_cache(tmp_path, monkeypatch):
    """chroma/loader.py:131-199: create_geometry_from_obj flattens, takes the BVH from the cache when
    it is there and otherwise builds and stores it.  The build itself needs the GPU library, so it is
    replaced here by the oracle's NumPy restatement of the same builder."""
    from chroma_lite_b200 import loader, demo
    from chroma_lite_b200.geometry import Geometry, Solid, Mesh
    from chroma_lite_b200.make import sphere
    from oracle import bvh_oracle
    built = []

    def build(mesh, target_degree=3):
        built.append(target_degree)
        o, sc, nodes, offs = bvh_oracle.make_recursive_grid_bvh(mesh.vertices, mesh.triangles, target_degree)
        return BVH(WorldCoords(o, sc), nodes.view(uint4)[:, 0], offs)
    monkeypatch.setattr(loader, 'make_recursive_grid_bvh', build)
    d = str(tmp_path / 'c')
    g1 = loader.create_geometry_from_obj(demo.tiny, cache_dir=d)              # callable -> Detector
    assert built == [3] and g1.bvh is not None and hasattr(g1, 'mesh') and hasattr(g1, 'num_channels')
    key = cache.mesh_hash(g1.mesh)
    assert cache.Cache(d).list_bvh(key) == ['default']
    g2 = loader.create_geometry_from_obj(demo.tiny(), cache_dir=d)            # cache hit: no build
    assert built == [3]
    assert np.array_equal(np.asarray(g2.bvh.nodes).view(np.uint32), np.asarray(g1.bvh.nodes).view(np.uint32))
    assert g2.bvh.layer_offsets == list(g1.bvh.layer_offsets)
    g3 = loader.create_geometry_from_obj(demo.tiny(), cache_dir=d, read_bvh_cache=False, update_bvh_cache=False)
    assert built == [3, 3] and len(g3.bvh) == len(g1.bvh)
    g4 = loader.create_geometry_from_obj(sphere(5.0, 8), cache_dir=d, auto_build_bvh=False)   # bare Mesh
    assert isinstance(g4, Geometry) and g4.bvh is None and len(g4.solids) == 1
    g5 = loader.create_geometry_from_obj(Solid(sphere(5.0, 8), scenes.optics.water, scenes.optics.water),
                                         cache_dir=d, update_bvh_cache=False)
    assert len(g5.bvh) > len(g5.mesh.triangles)
    assert cache.Cache(d).list_bvh(cache.mesh_hash(g5.mesh)) == []
    with pytest.raises(TypeError):
        loader.create_geometry_from_obj(42, cache_dir=d)
    g6 = demo.tiny()
    g6.bvh = 'kept'
    assert loader.create_geometry_from_obj(g6, cache_dir=d).bvh == 'kept'


def test_device_block_cache_reserve_and_reuse(monkeypatch):
    """gpuarray._Pool: blocks are cached by bucketed size; reserve() pre-populates a bucket (so that a pipeline
    never meets cudaMalloc while kernels run) and stops at the cache limit; take() hands reserved blocks out."""
    import ctypes as C
    from chroma_lite_b200 import gpuarray as ga

    class FakeLib(object):
        def __init__(self):
            self.next, self.freed = 0x1000, []

        def cb_malloc(self, size, pptr):
            pptr._obj.value = self.next
            self.next += int(size)
            return 0

        def cb_free(self, p):
            self.freed.append(p.value)
            return 0

    fake = FakeLib()
    monkeypatch.setattr(ga._lib, 'lib', lambda: fake)
    monkeypatch.setattr(ga._lib, '_lib', fake)
    pool = ga._Pool()
    assert pool.bucket(1) == 256 and pool.bucket(300) == 512 and pool.bucket(16 << 20) == 16 << 20
    assert pool.bucket((16 << 20) + 1) == 18 << 20                      # 2 MiB steps above 16 MiB
    assert pool.reserve(30000000, 3) == 3 and pool.reserve(30000000, 3) == 0 and pool.reserve(30000000, 5) == 2
    size = pool.bucket(30000000)
    assert len(pool.free[size]) == 5 and pool.cached == 5 * size
    got = [pool.take(size) for _ in range(6)]
    assert got[5] is None and len(set(got[:5])) == 5 and pool.cached == 0
    assert pool.give(got[0], size) and pool.take(size) == got[0]
    pool.LIMIT = 3 * size
    assert pool.reserve(30000000, 10) == 3                               # the cache limit bounds a reservation
    pool.release_all()
    assert len(fake.freed) == 3 and pool.cached == 0


def test_bench_warm_up_waits_for_stable_times():
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    series = iter([9.0, 5.0, 4.0, 3.0, 2.5] + [2.0] * 1000)
    n = bench.warm_until_stable(lambda: next(series), min_calls=10, min_s=0.0, max_s=5.0)
    assert n == 15                                                      # two consecutive groups of five within 2 %
    calls = []
    n = bench.warm_until_stable(lambda: calls.append(1) or float(len(calls)), min_calls=5, min_s=0.0, max_s=0.05)
    assert n == len(calls) and n >= 10                                   # never stable: stops at max_s
