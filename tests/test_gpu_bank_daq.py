"""Photon-bank utilities, DAQ, BVH build and the Simulation front end on the GPU."""
import numpy as np
import pytest

from chroma_lite_b200 import gpu, event, sim
from chroma_lite_b200 import gpuarray as ga
from chroma_lite_b200.bvh import make_recursive_grid_bvh
from oracle import orc, ref_driver, bvh_oracle
import scenes

pytestmark = pytest.mark.gpu


def propagated_bank(geo, n=120000, seed=3, max_steps=100):
    g = gpu.GPUDetector(geo)
    ph = scenes.point_source(n, seed=seed, wl_range=(300, 600))
    ph.evidx[:] = np.arange(n) % 3
    rng = gpu.get_rng_states(n, seed=21)
    gp = gpu.GPUPhotons(ph)
    gp.propagate(g, rng, max_blocks=(n + 255) // 256, max_steps=max_steps)
    return g, gp, rng


def test_select_matches_numpy_and_keeps_order(gpu_ready):
    geo = scenes.tiny_detector()
    g, gp, _ = propagated_bank(geo)
    host = gp.get()
    for flag in (event.SURFACE_DETECT, event.BULK_ABSORB | event.RAYLEIGH_SCATTER, 1 << 14):
        sel = gp.select(flag).get()
        mask = (host.flags & flag) != 0
        assert len(sel) == mask.sum()
        assert np.array_equal(sel.pos, host.pos[mask]) and np.array_equal(sel.flags, host.flags[mask])
        assert np.array_equal(sel.evidx, host.evidx[mask]) and np.array_equal(sel.t, host.t[mask])
    sub = gp.select(event.SURFACE_DETECT, start_photon=1000, nphotons=5000).get()
    m = (host.flags[1000:6000] & event.SURFACE_DETECT) != 0
    assert np.array_equal(sub.wavelengths, host.wavelengths[1000:6000][m])


def test_flat_hits_match_numpy(gpu_ready):
    geo = scenes.tiny_detector()
    g, gp, _ = propagated_bank(geo)
    host = gp.get()
    hits = gp.get_flat_hits(g)
    tri = host.last_hit_triangles
    det = (host.flags & event.SURFACE_DETECT) != 0
    chan = np.full(len(host), -1)
    ok = det & (tri > -1)
    chan[ok] = geo.solid_id_to_channel_index[geo.solid_id[tri[ok]]]
    mask = chan >= 0
    assert len(hits) == mask.sum() and len(hits) > 100
    assert np.array_equal(hits.channel.astype(np.int64), chan[mask])
    assert np.array_equal(hits.pos, host.pos[mask]) and np.array_equal(hits.evidx, host.evidx[mask])
    hm = gp.get_hits(g)
    assert sum(len(v) for v in hm.values()) == len(hits)


def test_duplicate_and_iterate_copies(gpu_ready):
    ph = scenes.point_source(1000, seed=8)
    gp = gpu.GPUPhotons(ph, ncopies=3)
    host = gp.get()
    assert len(host) == 3000
    for c in range(3):
        assert np.array_equal(host.pos[c * 1000:(c + 1) * 1000], ph.pos)
        assert np.array_equal(host.wavelengths[c * 1000:(c + 1) * 1000], ph.wavelengths)
    copies = list(gp.iterate_copies())
    assert len(copies) == 3 and np.array_equal(copies[2].get().dir, ph.dir)


def test_copy_queue(gpu_ready):
    ph = scenes.point_source(5000, seed=9)
    gp = gpu.GPUPhotons(ph)
    q = np.random.default_rng(0).permutation(5000)[:777].astype(np.uint32)
    out = gp.copy_queue(ga.to_gpu(q), len(q)).get()
    assert np.array_equal(out.pos, ph.pos[q]) and np.array_equal(out.pol, ph.pol[q])


def test_daq_vs_reference_kernel_and_oracle(gpu_ready):
    geo = scenes.tiny_detector()
    g, gp, _ = propagated_bank(geo, n=150000)
    host = gp.get()
    n = len(host)
    # engine
    rng = gpu.get_rng_states(n, seed=5)
    daq = gpu.GPUDaq(g)
    daq.begin_acquire()
    daq.acquire(gp, rng, nthreads_per_block=64, max_blocks=(n + 63) // 64)
    ch = daq.end_acquire().get()
    # reference run_daq on the same photons and seed
    desc, keep = scenes.desc_of(geo)
    rg = ref_driver.RefGeometry(desc, keep)
    rg.attach_detector(geo)
    rrng = ref_driver.RefRNG(n, seed=5)
    rt, rq, rh, rti, rqi = ref_driver.run_daq(rg, ref_driver.RefPhotons(host), rrng, nthreads_per_block=64,
                                              max_blocks=(n + 63) // 64)
    assert np.array_equal(ch.flags, rh)
    assert np.array_equal(ch.q, rq)
    assert np.array_equal(ch.t, rt)
    assert ch.hit.sum() > 20 and np.array_equal(ch.hit, rt < 1e8)
    # CPU oracle (integer accumulators exact, times to float tolerance)
    tint, qint, hist, unit = orc.run_daq(orc.HostBank(host), orc.rng_init(5, 0, n), geo, geo.solid_id)
    assert np.array_equal(hist, ch.flags)
    # charge/charge_unit uses an approximate division on the GPU: +-1 count per hit at most
    dq = np.abs(qint.astype(np.int64) - daq.channel_q_int_gpu.get().astype(np.int64))
    assert dq.max() <= 3 and (dq == 0).mean() > 0.8
    assert np.allclose(tint.view(np.float32), ch.t, rtol=1e-6)


def test_daq_many_vs_reference_kernel(gpu_ready):
    # ndaq > 1: run_daq_many, one block per photon, threads striding over the virtual DAQ copies,
    # curand_normal (Box-Muller with its cached second value) in the time smearing (daq.cu:88-150)
    geo = scenes.tiny_detector()
    g, gp, _ = propagated_bank(geo, n=60000)
    host = gp.get()
    n = len(host)
    ndaq, tpb, max_blocks = 6, 64, 1024
    pool = tpb * max_blocks
    rng = gpu.get_rng_states(pool, seed=11)
    daq = gpu.GPUDaq(g, ndaq=ndaq)
    daq.begin_acquire()
    daq.acquire(gp, rng, nthreads_per_block=tpb, max_blocks=max_blocks, weight=0.8)
    chans = daq.end_acquire()
    desc, keep = scenes.desc_of(geo)
    rg = ref_driver.RefGeometry(desc, keep)
    rg.attach_detector(geo)
    rrng = ref_driver.RefRNG(pool, seed=11)
    rt, rq, rh, rti, rqi = ref_driver.run_daq_many(rg, ref_driver.RefPhotons(host), rrng, ndaq, nthreads_per_block=tpb,
                                                    max_blocks=max_blocks, weight=0.8)
    assert np.array_equal(daq.channel_history_gpu.get(), rh)
    assert np.array_equal(daq.channel_q_int_gpu.get(), rqi)
    assert np.array_equal(daq.earliest_time_int_gpu.get(), rti)
    assert np.array_equal(daq.earliest_time_gpu.get(), rt)
    assert (rti != np.float32(1e9).view(np.uint32)).sum() > 10 * ndaq          # every copy saw hits
    # the copies are independent draws: they differ from each other
    t = daq.earliest_time_gpu.get().reshape(ndaq, -1)
    assert not np.array_equal(t[0], t[1])
    # RNG pool advanced identically (incl. the lanes that never drew)
    assert np.array_equal(rng.get(), rrng.states6())
    copies = list(chans.iterate_copies())
    assert len(copies) == ndaq


def test_daq_time_and_charge_response(gpu_ready):
    # test/test_detector.py: time spread and charge mean/rms of single photoelectrons
    from chroma_lite_b200.detector import Detector
    from chroma_lite_b200.geometry import Solid, vacuum
    from chroma_lite_b200.make import cube
    from chroma_lite_b200.demo import optics
    det = Detector(vacuum)
    cath = scenes.Surface('cathode')
    cath.set('detect', 1.0)
    det.add_pmt(Solid(cube(100.0), vacuum, vacuum, surface=cath))
    det.set_time_dist_gaussian(1.2, -6.0, 6.0)
    det.set_charge_dist_gaussian(1.0, 0.1, 0.5, 1.5)
    scenes.with_bvh(det)
    g = gpu.GPUDetector(det)
    daq = gpu.GPUDaq(g)
    n = 2000
    rng = gpu.get_rng_states(4096, seed=3)
    ts, qs = [], []
    for i in range(n):
        # t0 = 100 ns: the DAQ orders times by raw float bits, valid for t >= 0 only
        ph = event.Photons(np.zeros((1, 3)), np.array([[1.0, 0, 0]]), np.array([[0, 1.0, 0]]), np.array([400.0]),
                           t=np.array([100.0]))
        gp = gpu.GPUPhotons(ph)
        gp.propagate(g, rng, max_steps=10)
        daq.begin_acquire()
        daq.acquire(gp, rng)
        ch = daq.end_acquire().get()
        if ch.hit[0]:            # a photon can be absorbed in the bulk (P ~ 1e-4)
            ts.append(ch.t[0])
            qs.append(ch.q[0])
    ts, qs = np.array(ts), np.array(qs)
    assert len(ts) > 0.99 * n
    assert abs(ts.std() - 1.2) < 0.1 and abs(ts.mean() - 100.0 - 50.0 / 299.792458) < 0.15
    assert abs(qs.mean() - 1.0) < 0.02 and abs(qs.std() - 0.1) < 0.02


def test_bvh_build_matches_reference_kernels(gpu_ready):
    # The reference's builder with ALL FOUR of its device kernels run from oracle/_ref/bvh.cubin
    # (make_leaves, make_parents_detailed, copy_and_offset, collapse_child: chroma/cuda/bvh.cu:148, 269,
    # 364, 530, launched as chroma/gpu/bvh.py:18-130, 239-267 does) and the host grouping of
    # chroma/bvh/grid.py:11-95 restated in NumPy: the engine's builder must give the identical tree,
    # node for node.
    for geo in (scenes.sphere_scene(16), scenes.tiny_detector(), scenes.scintillator_scene(12), scenes.ref_tiny_detector()):
        bvh = make_recursive_grid_bvh(geo.mesh)
        o, s, nodes, offs = bvh_oracle.make_recursive_grid_bvh(
            geo.mesh.vertices, geo.mesh.triangles, leaves=ref_driver.make_leaves, parents=ref_driver.merge_nodes_detailed,
            concatenate=ref_driver.concatenate_layers, collapse=ref_driver.collapse_chains)
        # the C restatement of the three small kernels agrees with the reference's kernels too
        o1, s1, nodes1, offs1 = bvh_oracle.make_recursive_grid_bvh(geo.mesh.vertices, geo.mesh.triangles,
                                                                   leaves=ref_driver.make_leaves)
        assert np.array_equal(nodes1, nodes) and list(offs1) == list(offs)
        assert np.array_equal(bvh.world_coords.world_origin, o) and bvh.world_coords.world_scale == s
        mine = bvh.nodes.view(np.uint32).reshape(-1, 4)
        assert list(bvh.layer_offsets) == list(offs)
        assert np.array_equal(mine, nodes)
        # CPU restatement (exact division instead of the device's approximate one): same
        # topology unless a vertex sits exactly on a quantum boundary
        o2, s2, nodes2, offs2 = bvh_oracle.make_recursive_grid_bvh(geo.mesh.vertices, geo.mesh.triangles)
        if nodes2.shape == nodes.shape:
            assert (np.abs(nodes2[:, :3].astype(np.int64) - nodes[:, :3].astype(np.int64)) % 65536 <= 1).mean() > 0.9


def test_loader_builds_caches_and_reloads(gpu_ready, tmp_path):
    """chroma/loader.py:162-199 end to end: build on the GPU, store, reload (memory-mapped) and propagate
    with the reloaded tree -- same photons as with the freshly built one."""
    from chroma_lite_b200 import loader, demo
    d = str(tmp_path / 'cache')
    g1 = loader.create_geometry_from_obj(demo.tiny, cache_dir=d)
    g2 = loader.create_geometry_from_obj(demo.tiny, cache_dir=d)
    assert np.array_equal(np.asarray(g2.bvh.nodes).view(np.uint32), g1.bvh.nodes.view(np.uint32))
    ends = []
    for g in (g1, g2):
        gp = gpu.GPUPhotons(scenes.point_source(20000, seed=2, wl_range=(300, 600)))
        gp.propagate(gpu.GPUDetector(g), gpu.get_rng_states(20000, seed=5), max_blocks=128, max_steps=20)
        ends.append(gp.get())
    assert np.array_equal(ends[0].flags, ends[1].flags) and np.array_equal(ends[0].pos, ends[1].pos)


def test_prepared_tree_cache_is_shared_and_exact(gpu_ready, tmp_path, monkeypatch):
    """CHROMA_B200_TREE_CACHE: what cb_geometry_create derives on the host (reference test rank, leaf boxes, the
    engine's traversal tree) is written once and read back by later creations (other ranks of a node): same
    hits, same propagation, and a second tree option gets its own file."""
    import os
    geo = scenes.ref_tiny_detector()
    ph = scenes.point_source(30000, seed=4, wl_range=(300, 600))

    def run():
        g = gpu.GPUDetector(geo)
        gp = gpu.GPUPhotons(ph)
        gp.propagate(g, gpu.get_rng_states(len(ph), seed=8), nthreads_per_block=256, max_blocks=(len(ph) + 255) // 256, max_steps=50)
        tri, dist = gpu.intersect_mesh(g, ph.pos, ph.dir)
        return gp.get(), tri.get(), dist.get()
    monkeypatch.delenv('CHROMA_B200_TREE_CACHE', raising=False)
    base = run()
    d = tmp_path / 'trees'
    monkeypatch.setenv('CHROMA_B200_TREE_CACHE', str(d))
    first = run()
    files = sorted(os.listdir(str(d)))
    assert len(files) == 1 and files[0].startswith('tree_') and files[0].endswith('.bin')
    stamp = os.path.getmtime(str(d / files[0]))
    second = run()                                  # served from the file
    assert os.path.getmtime(str(d / files[0])) == stamp and sorted(os.listdir(str(d))) == files
    for out in (first, second):
        assert np.array_equal(out[1], base[1]) and np.array_equal(out[2].view(np.uint32), base[2].view(np.uint32))
        for f in ('pos', 't', 'flags', 'last_hit_triangles'):
            assert np.array_equal(getattr(out[0], f), getattr(base[0], f)), f
    monkeypatch.setenv('CHROMA_B200_TREE', 'solids')
    other = run()
    assert len(os.listdir(str(d))) == 2             # keyed by the tree option too
    assert np.array_equal(other[1], base[1])


def test_bvh_is_conservative(gpu_ready):
    geo = scenes.tiny_detector()
    bvh = make_recursive_grid_bvh(geo.mesh)
    from chroma_lite_b200.bvh import unpack_nodes
    u = unpack_nodes(bvh.nodes)
    leaves = u[u['nchild'] == 0]
    leaves = leaves[leaves['child'] < len(geo.mesh.triangles)]
    tri = geo.mesh.assemble()[leaves['child'].astype(np.int64)]
    wc = bvh.world_coords
    for a, (lo, hi) in enumerate((('xlo', 'xhi'), ('ylo', 'yhi'), ('zlo', 'zhi'))):
        lo_w = wc.world_origin[a] + leaves[lo].astype(np.float64) * wc.world_scale
        hi_w = wc.world_origin[a] + leaves[hi].astype(np.float64) * wc.world_scale
        assert (lo_w <= tri[:, :, a].min(axis=1) + 1e-3).all() and (hi_w >= tri[:, :, a].max(axis=1) - 1e-3).all()


def test_simulation_front_end(gpu_ready):
    geo = scenes.tiny_detector()
    s = sim.Simulation(geo, seed=12, nthreads_per_block=256, max_blocks=512)
    evs = [scenes.point_source(20000, seed=k, wl_range=(300, 600)) for k in range(3)]
    out = list(s.simulate(evs, keep_photons_end=True, keep_hits=True, keep_flat_hits=True, run_daq=True,
                          max_steps=100, photons_per_batch=45000))
    assert len(out) == 3
    for ev in out:
        assert len(ev.photons_end) == 20000 and ev.photons_beg is None
        assert ((ev.photons_end.flags & event.TERMINAL_MASK) != 0).mean() > 0.99
        assert len(ev.flat_hits) > 10 and sum(len(v) for v in ev.hits.values()) == len(ev.flat_hits)
        assert ev.channels.hit.sum() > 5 and (ev.channels.t[ev.channels.hit] < 1e8).all()
        hit_channels = np.unique(ev.flat_hits.channel)
        assert set(np.flatnonzero(ev.channels.hit)) <= set(hit_channels.astype(int))
    single = list(s.simulate(scenes.point_source(1000, seed=5), keep_photons_end=True, max_steps=10))
    assert len(single) == 1 and len(single[0].photons_end) == 1000


def test_pipeline_defers_collection_without_changing_results(gpu_ready):
    """simulate() only enqueues hit compaction and DAQ behind the propagate kernels and collects them
    from the consumer thread while later batches propagate; the events must be exactly those of the
    synchronous batch-by-batch path (_simulate_batch: the reference's order, chroma/sim.py:54-154),
    `gpu_daq` must hold the yielded event's arrays (what a caller folds into run-level accumulators),
    and more events per batch than ASYNC_EVENTS_MAX must take the synchronous read-back."""
    geo = scenes.tiny_detector()
    evs = [scenes.point_source(7000 + 300 * k, seed=70 + k, wl_range=(300, 600)) for k in range(7)]
    kw = dict(keep_photons_end=True, keep_hits=False, keep_flat_hits=True, run_daq=True, max_steps=60)
    fields = ('pos', 'dir', 'pol', 'wavelengths', 't', 'last_hit_triangles', 'flags', 'weights', 'evidx', 'channel')
    for per_batch in (1, 15000, 10 ** 9):
        s_sync = sim.Simulation(geo, seed=21, nthreads_per_block=256, max_blocks=512)
        ref, batch, n = [], [], 0
        for k, p in enumerate(evs):                       # the same batching rule as simulate()
            ev = event.Event(photons_beg=p)
            p.evidx[:] = len(batch)
            batch.append(ev)
            n += len(p)
            if n >= per_batch or k == len(evs) - 1:
                ref += list(s_sync._simulate_batch(batch, **kw))
                batch, n = [], 0
        s = sim.Simulation(geo, seed=21, nthreads_per_block=256, max_blocks=512)
        s.ASYNC_EVENTS_MAX = 3                              # the one-batch case (7 events) reads back synchronously
        run = gpu.GPUDaq(s.gpu_geometry)
        run.begin_acquire()
        out, q_sum = [], 0
        for ev in s.simulate([event.Event(photons_beg=p) for p in evs], photons_per_batch=per_batch, **kw):
            if s.last_timings.get('deferred'):
                run.fold(s.gpu_daq, wait=False)
                q_sum = q_sum + s.gpu_daq.channel_q_int_gpu.get().astype(np.int64)
            out.append(ev)
        assert len(out) == len(ref) == 7
        assert s.last_timings['deferred'] == (per_batch != 10 ** 9)
        for a, b in zip(ref, out):
            assert len(b.flat_hits) > 5
            for f in fields:
                assert np.array_equal(getattr(a.flat_hits, f), getattr(b.flat_hits, f)), (per_batch, f)
            assert np.array_equal(a.photons_end.flags, b.photons_end.flags)
            assert np.array_equal(a.channels.t, b.channels.t) and np.array_equal(a.channels.q, b.channels.q)
            assert np.array_equal(a.channels.flags, b.channels.flags) and np.array_equal(a.channels.hit, b.channels.hit)
        if per_batch != 10 ** 9:
            run.end_acquire()
            assert np.array_equal(run.channel_q_int_gpu.get().astype(np.int64), q_sum)
            assert np.array_equal(run.earliest_time_gpu.get(), np.min([e.channels.t for e in out], axis=0))
    # an event without any hit, and an empty selection, through the deferred path
    s = sim.Simulation(geo, seed=3, nthreads_per_block=256, max_blocks=512)
    dark = scenes.point_source(500, seed=1, wl_range=(300, 600))
    dark.flags[:] = event.SURFACE_ABSORB                    # terminal before the first step: nothing propagates
    ev = next(s.simulate(dark, keep_flat_hits=True, keep_hits=True, run_daq=True, max_steps=5))
    assert len(ev.flat_hits) == 0 and ev.hits == {} and not ev.channels.hit.any()


def _gpu_view(gp):
    from types import SimpleNamespace
    return SimpleNamespace(pos=gp.pos, dir=gp.dir, pol=gp.pol, wavelengths=gp.wavelengths, t=gp.t,
                           last_hit_triangles=gp.last_hit_triangles, flags=gp.flags, weights=gp.weights,
                           evidx=gp.evidx)


def test_gpu_photons_from_device_arrays(gpu_ready):
    """The reference's test/test_gpu_photon_gpu_input.py:50-85: a bank built from device arrays holds the
    same photons (the constructor copies device-to-device, gpu/photon.py:62-89), replicates them with
    ncopies, and resets the fields it is told not to copy."""
    ph = scenes.point_source(777, seed=4, wl_range=(300, 600))
    ph.flags[:] = np.arange(777) % 2
    ph.weights[:] = 0.5
    ph.last_hit_triangles[:] = 7
    src = gpu.GPUPhotons(ph)
    same = gpu.GPUPhotons(_gpu_view(src))
    assert same.true_nphotons == 777 and len(same) == 777
    a, b = src.get(), same.get()
    for f in ('pos', 'dir', 'pol', 'wavelengths', 't', 'last_hit_triangles', 'flags', 'weights', 'evidx'):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    dupe = gpu.GPUPhotons(_gpu_view(src), ncopies=2)
    d = dupe.get()
    assert len(dupe.pos) == 2 * 777
    assert np.array_equal(d.pos[:777], d.pos[777:]) and np.array_equal(d.flags[:777], d.flags[777:])
    assert np.array_equal(d.pos[:777], a.pos)
    reset = gpu.GPUPhotons(_gpu_view(src), copy_flags=False, copy_triangles=False, copy_weights=False)
    assert int(reset.flags.gpudata) != int(src.flags.gpudata)
    assert (reset.flags.get() == 0).all() and (reset.last_hit_triangles.get() == -1).all()
    assert np.allclose(reset.weights.get(), 1.0)


def test_bank_upload_from_host_arrays_in_one_call(gpu_ready):
    """GPUPhotons from host arrays goes through cb_photon_bank_upload: what is copied arrives bit for bit
    (any input dtype / layout, converted like the reference's to_gpu calls), what is not copied gets the
    constructor's defaults (gpu/photon.py:46-62), for every copy of ncopies."""
    ph = scenes.point_source(300001, seed=8, wl_range=(300, 600))
    ph.flags[:] = np.arange(len(ph)) % 5
    ph.weights[:] = 0.25
    ph.last_hit_triangles[:] = np.arange(len(ph)) % 11
    ph.evidx[:] = 9
    full = gpu.GPUPhotons(ph).get()
    for f in ('pos', 'dir', 'pol', 'wavelengths', 't', 'last_hit_triangles', 'flags', 'weights', 'evidx'):
        assert np.array_equal(getattr(full, f), getattr(ph, f)), f
    # all-zero flags / times are zeroed on the device instead of being sent; anything else is sent
    ph.flags[:] = 0
    ph.t[:] = 1.5
    sent = gpu.GPUPhotons(ph)
    assert sent.h2d_bytes == len(ph) * (36 + 4 * 5)
    back = sent.get()
    assert (back.flags == 0).all() and (back.t == 1.5).all() and np.array_equal(back.last_hit_triangles, ph.last_hit_triangles)
    ph.t[:] = 0
    ph.t[7] = -0.0                                         # not the all-zero bit pattern: sent as it is
    assert gpu.GPUPhotons(ph).h2d_bytes == len(ph) * (36 + 4 * 5)
    ph.t[7] = 0
    none = gpu.GPUPhotons(ph)
    assert none.h2d_bytes == len(ph) * (36 + 4 * 4) and (none.get().t == 0).all()
    odd = event.Photons(np.asfortranarray(ph.pos.astype(np.float64)), ph.dir, ph.pol, ph.wavelengths.astype(np.float64), ph.t,
                        ph.last_hit_triangles.astype(np.int64), ph.flags, ph.weights, ph.evidx)
    got = gpu.GPUPhotons(odd, ncopies=2, copy_flags=False, copy_triangles=False, copy_weights=False, evidx_value=3).get()
    assert len(got) == 2 * len(ph)
    for half in (slice(0, len(ph)), slice(len(ph), None)):
        assert np.array_equal(got.pos[half], ph.pos) and np.array_equal(got.wavelengths[half], ph.wavelengths)
        assert (got.flags[half] == 0).all() and (got.last_hit_triangles[half] == -1).all()
        assert (got.weights[half] == 1.0).all() and (got.evidx[half] == 3).all()
    with pytest.raises(ValueError):
        gpu.GPUPhotons(event.Photons(ph.pos, ph.dir, ph.pol, ph.wavelengths[:-1], ph.t))


def test_simulate_accepts_gpu_photons(gpu_ready, monkeypatch):
    """test_gpu_photon_gpu_input.py:87-108: device-resident sources never go through the host join; several
    of them in one batch are stacked on the device (sim.py:170-226) and give the events of the host path."""
    geo = scenes.tiny_detector()
    evs = [scenes.point_source(6000 + 500 * k, seed=40 + k, wl_range=(300, 600)) for k in range(3)]
    kw = dict(keep_photons_end=True, keep_flat_hits=True, keep_hits=False, run_daq=True, max_steps=50,
              photons_per_batch=13000)
    host = list(sim.Simulation(geo, seed=9, nthreads_per_block=256, max_blocks=128).simulate(
        [event.Event(photons_beg=p) for p in evs], **kw))
    s = sim.Simulation(geo, seed=9, nthreads_per_block=256, max_blocks=128)
    banks = [gpu.GPUPhotons(p) for p in evs]

    def no_join(*a, **k):
        raise AssertionError('CPU join should not be used for GPU sources')
    monkeypatch.setattr(event.Photons, 'join', staticmethod(no_join))
    dev = list(s.simulate([event.Event(photons_beg=b) for b in banks], **kw))
    monkeypatch.undo()
    assert len(dev) == len(host) == 3
    for h, d in zip(host, dev):
        assert np.array_equal(h.photons_end.flags, d.photons_end.flags)
        assert np.array_equal(h.photons_end.pos, d.photons_end.pos)
        assert np.array_equal(h.flat_hits.channel, d.flat_hits.channel) and len(d.flat_hits) > 5
        assert np.array_equal(h.channels.t, d.channels.t) and np.array_equal(h.channels.q, d.channels.q)
    stacked = sim.Simulation._stack_gpu_photon_sources([_gpu_view(b) for b in banks])
    assert stacked.true_nphotons == sum(len(p) for p in evs)
    assert np.array_equal(stacked.t.get(), np.concatenate([p.t for p in evs]))
    assert sim.Simulation._stack_gpu_photon_sources([_gpu_view(banks[0]), evs[1]]) is None
    one = list(s.simulate([event.Event(photons_beg=gpu.GPUPhotons(evs[0][:1]))], keep_hits=False,
                          keep_flat_hits=False, run_daq=False, max_steps=1))
    assert len(one) == 1


def test_errors_are_loud(gpu_ready):
    from chroma_lite_b200 import _lib
    geo = scenes.water_box(10.0)
    g = gpu.GPUGeometry(geo)
    with pytest.raises(AssertionError):
        gpu.GPUDaq(gpu.GPUDetector(scenes.with_bvh(__import__('chroma_lite_b200').detector.Detector(None))) if False else type('X', (), {'nchannels': 0})())
    gp = gpu.GPUPhotons(scenes.point_source(10))
    with pytest.raises(_lib.ChromaB200Error):
        gp.propagate(g, type('R', (), {'handle': 987654})(), max_steps=1)


def test_rat_bridge_request(gpu_ready):
    """bin/chroma-server-rat's request -> propagate -> reply cycle through wire.handle_rat_request: the
    reply holds the flat hits of the same event, grouped by channel."""
    from chroma_lite_b200 import sim, wire
    det = scenes.tiny_detector()
    ph = scenes.point_source(80000, seed=31, wl_range=(300, 600))
    msg = wire.encode_rat_request(ph, event_id=12)
    s1 = sim.Simulation(det, seed=3, nthreads_per_block=256, max_blocks=512)
    hits, evid = wire.decode_rat_reply(wire.handle_rat_request(s1, msg, max_steps=100))
    s2 = sim.Simulation(det, seed=3, nthreads_per_block=256, max_blocks=512)
    ev = next(s2.simulate(ph, keep_flat_hits=True, keep_hits=False, max_steps=100))
    order = np.argsort(ev.flat_hits.channel, kind='stable')
    assert evid == 12 and len(hits) == len(ev.flat_hits) > 100
    assert np.array_equal(hits.channel, ev.flat_hits.channel[order])
    assert np.array_equal(hits.t, ev.flat_hits.t[order]) and np.array_equal(hits.pos, ev.flat_hits.pos[order])
    empty, _ = wire.decode_rat_reply(wire.handle_rat_request(s1, wire.encode_rat_request(ph[:0], event_id=1)))
    assert len(empty) == 0


def test_photon_server_request(gpu_ready):
    """bin/chroma-server:31-40: a pickled Photons object in, the pickled event with photons_end out."""
    import pickle
    from chroma_lite_b200 import sim, wire
    det = scenes.tiny_detector()
    ph = scenes.point_source(30000, seed=17, wl_range=(300, 600))
    s1 = sim.Simulation(det, seed=4, nthreads_per_block=256, max_blocks=128)
    ev = pickle.loads(wire.handle_photons_request(s1, pickle.dumps(ph), max_steps=100))
    s2 = sim.Simulation(det, seed=4, nthreads_per_block=256, max_blocks=128)
    ref = next(s2.simulate(ph, keep_photons_end=True, max_steps=100))
    assert len(ev.photons_end) == 30000 and ev.photons_beg is None
    assert np.array_equal(ev.photons_end.flags, ref.photons_end.flags)
    assert np.array_equal(ev.photons_end.pos, ref.photons_end.pos)
    assert np.array_equal(ev.flat_hits.channel, ref.flat_hits.channel)

