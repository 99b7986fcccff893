"""Acceptance at BASELINE.json's full sizes (north_star's correctness statement):

* config 3: 2.5 M photons in the 29k-PMT detector (36.96 M triangles), engine vs the
  reference's own propagate kernel on the same inputs and seed: >= 99.9 % of the photons
  with identical flags and last_hit_triangle, position and time within 1e-4 relative;
  per-channel hit counts and hit-time distributions of an INDEPENDENTLY seeded engine
  run statistically consistent with the reference's (chi2 / KS, p > 0.01);
* size-independent properties at that size: results do not depend on how the scheduler
  splits the work (wavefront steps vs persistent tail), every detected photon sits on a
  triangle, hits + DAQ agree with a NumPy evaluation of the same bank;
* config 2: 10 M random rays against a 1.18 M-triangle mesh, triangle and distance
  bit-exact vs the reference's intersect_mesh.

The detector is built once per module (bench.py's builder and cache).
"""
import numpy as np
import pytest
from scipy import stats

import bench
from chroma_lite_b200 import gpu, event
from chroma_lite_b200.gpu.geometry import make_desc
from oracle import ref_driver

pytestmark = pytest.mark.gpu
TERM = event.TERMINAL_MASK
N = 2500000
MAX_STEPS = 100


@pytest.fixture(scope='module')
def full(gpu_ready):
    det = bench.build_detector('pmt29k', {})
    g = gpu.GPUDetector(det)
    desc, keep = make_desc(det)
    rg = ref_driver.RefGeometry(desc, keep)
    rg.attach_detector(det)
    return {'det': det, 'g': g, 'rg': rg, 'runs': {}}


def engine_run(full, seed, event_seed=1000):
    key = (seed, event_seed)
    if key not in full['runs']:
        ph = bench.make_event(N, seed=event_seed)
        rng = gpu.get_rng_states(N, seed=seed)
        gp = gpu.GPUPhotons(ph)
        gp.propagate(full['g'], rng, nthreads_per_block=512, max_blocks=(N + 511) // 512, max_steps=MAX_STEPS)
        full['runs'][key] = (gp, gp.get(), rng.get())
    return full['runs'][key]


def reference_run(full, seed, event_seed=1000):
    key = ('ref', seed, event_seed)
    if key not in full['runs']:
        ph = bench.make_event(N, seed=event_seed)
        rng = ref_driver.RefRNG(N, seed=seed)
        rp = ref_driver.RefPhotons(ph)
        rp.propagate(full['rg'], rng, nthreads_per_block=256, max_steps=MAX_STEPS, force_single_launch=True)
        full['runs'][key] = (rp, rp.get(), rng.states6())
    return full['runs'][key]


def test_config3_full_size_vs_reference_kernel(full):
    _, mine, st_mine = engine_run(full, 42)
    _, ref, st_ref = reference_run(full, 42)
    same = (mine.flags == ref.flags) & (mine.last_hit_triangles == ref.last_hit_triangles)
    assert same.mean() >= 0.999, 'identical histories: %.5f' % same.mean()
    scale = np.maximum(np.abs(ref.pos[same]).max(axis=1), 1.0)
    assert (np.abs(mine.pos[same] - ref.pos[same]).max(axis=1) / scale < 1e-4).mean() > 0.9999
    assert np.isclose(mine.t[same], ref.t[same], rtol=1e-4, atol=1e-3).mean() > 0.9999
    assert (st_mine[same] == st_ref[same]).all(axis=1).mean() > 0.9999      # RNG streams advanced identically
    # every photon ended, and the flag mix is the detector's (~13 % PMT coverage)
    assert ((mine.flags & TERM) != 0).mean() > 0.9999
    det = ((mine.flags & event.SURFACE_DETECT) != 0)
    assert 0.005 < det.mean() < 0.05
    assert (mine.last_hit_triangles[det] >= 0).all()
    print('full-size identical fraction %.6f, detected %.4f' % (same.mean(), det.mean()))


def channel_of(full, bank):
    """NumPy evaluation of get_flat_hits' selection (gpu/photon.py:141-209)."""
    det = full['det']
    m = ((bank.flags & event.SURFACE_DETECT) != 0) & (bank.last_hit_triangles >= 0)
    ch = np.full(len(bank.flags), -1, dtype=np.int64)
    ch[m] = np.asarray(det.solid_id_to_channel_index)[det.solid_id[bank.last_hit_triangles[m]]]
    return ch


def test_config3_statistics_independent_seed(full):
    """A differently seeded engine run against the reference run: hit counts per group of
    channels (chi2) and the hit-time distribution (KS) agree at p > 0.01."""
    _, mine, _ = engine_run(full, 4242)
    _, ref, _ = reference_run(full, 42)
    cm, cr = channel_of(full, mine), channel_of(full, ref)
    hm, hr = cm >= 0, cr >= 0
    assert abs(hm.sum() - hr.sum()) < 5 * np.sqrt(hr.sum())
    groups = 64
    a = np.bincount(cm[hm] % groups, minlength=groups).astype(float)
    b = np.bincount(cr[hr] % groups, minlength=groups).astype(float)
    chi2, p, _, _ = stats.chi2_contingency(np.vstack([a, b]))
    assert p > 0.01, 'per-channel-group hit counts: chi2 p = %.4g' % p
    ks = stats.ks_2samp(mine.t[hm], ref.t[hr])
    assert ks.pvalue > 0.01, 'hit times: KS p = %.4g' % ks.pvalue
    for bit in (event.SURFACE_DETECT, event.SURFACE_ABSORB, event.BULK_ABSORB, event.RAYLEIGH_SCATTER,
                event.REFLECT_DIFFUSE, event.REFLECT_SPECULAR):
        fa, fb = ((mine.flags & bit) != 0).mean(), ((ref.flags & bit) != 0).mean()
        assert abs(fa - fb) < 5 * np.sqrt(max(fb, 1e-6) / N) + 1e-5, (hex(bit), fa, fb)
    print('chi2 p %.3f, KS p %.3f, hits %d vs %d' % (p, ks.pvalue, hm.sum(), hr.sum()))


def test_config3_scheduler_invariance_full_size(full, monkeypatch):
    """Same seed, different schedules (all wavefront steps down to 20 k photons; tail from the
    second step on): bit-identical banks and RNG pools."""
    _, base, st_base = engine_run(full, 42)
    for tail in ('20000', '1000000'):
        monkeypatch.setenv('CHROMA_B200_TAIL', tail)
        ph = bench.make_event(N, seed=1000)
        rng = gpu.get_rng_states(N, seed=42)
        gp = gpu.GPUPhotons(ph)
        gp.propagate(full['g'], rng, nthreads_per_block=512, max_blocks=(N + 511) // 512, max_steps=MAX_STEPS)
        out = gp.get()
        for f in ('pos', 'dir', 'pol', 'wavelengths', 't', 'last_hit_triangles', 'flags', 'weights'):
            assert np.array_equal(getattr(out, f).view(np.uint32), getattr(base, f).view(np.uint32)), (tail, f)
        assert np.array_equal(rng.get(), st_base)
    monkeypatch.delenv('CHROMA_B200_TAIL')


def test_config3_hits_and_daq_full_size(full):
    gp, mine, _ = engine_run(full, 42)
    g = full['g']
    ch = channel_of(full, mine)
    hits = gp.get_flat_hits(g)
    sel = np.flatnonzero(ch >= 0)
    assert len(hits) == len(sel)
    assert np.array_equal(hits.channel.astype(np.int64), ch[sel])            # order preserved
    assert np.array_equal(hits.t.view(np.uint32), mine.t[sel].view(np.uint32))
    # DAQ: a channel is hit iff some selected photon maps to it (weight 1: every such photon counts)
    rng = gpu.get_rng_states(N, seed=7)
    daq = gpu.GPUDaq(g)
    daq.begin_acquire()
    daq.acquire(gp, rng, nthreads_per_block=512, max_blocks=(N + 511) // 512)
    chans = daq.end_acquire().get()
    expect = np.zeros(len(chans.hit), dtype=bool)
    expect[np.unique(ch[sel])] = True
    assert np.array_equal(chans.hit, expect)
    # and the reference's run_daq on the same bank and seed gives the same channels bit for bit
    rrng = ref_driver.RefRNG(N, seed=7)
    rt, rq, rh, _, _ = ref_driver.run_daq(full['rg'], ref_driver.RefPhotons(mine), rrng, nthreads_per_block=512,
                                          max_blocks=(N + 511) // 512)
    assert np.array_equal(chans.t, rt) and np.array_equal(chans.q, rq) and np.array_equal(chans.flags, rh)


def test_config4_ten_million_photons_scintillator(gpu_ready):
    """BASELINE config 4 at its full 10 M photons: scintillator + WLS re-emission, dichroic, angular and
    thin-film surfaces, engine vs the reference kernel on the same inputs and seed."""
    import scenes
    geo = scenes.scintillator_scene(96)
    n = 10000000
    ph = scenes.point_source(n, seed=21, wl_range=(250, 450))
    g = gpu.GPUDetector(geo)
    rng = gpu.get_rng_states(n, seed=9)
    gp = gpu.GPUPhotons(ph)
    gp.propagate(g, rng, nthreads_per_block=512, max_blocks=(n + 511) // 512, max_steps=200)
    mine = gp.get()
    desc, keep = make_desc(geo)
    rg = ref_driver.RefGeometry(desc, keep)
    rrng = ref_driver.RefRNG(n, seed=9)
    rp = ref_driver.RefPhotons(ph)
    rp.propagate(rg, rrng, nthreads_per_block=256, max_steps=200, force_single_launch=True)
    ref = rp.get()
    same = (mine.flags == ref.flags) & (mine.last_hit_triangles == ref.last_hit_triangles)
    assert same.mean() >= 0.995, 'identical histories: %.5f' % same.mean()
    scale = np.maximum(np.abs(ref.pos[same]).max(axis=1), 1.0)
    assert (np.abs(mine.pos[same] - ref.pos[same]).max(axis=1) / scale < 1e-4).mean() > 0.999
    assert np.isclose(mine.t[same], ref.t[same], rtol=1e-4, atol=1e-3).mean() > 0.999
    assert ((mine.flags & TERM) != 0).mean() > 0.999
    for bit in (event.BULK_REEMIT, event.SURFACE_REEMIT, event.SURFACE_TRANSMIT, event.SURFACE_DETECT,
                event.RAYLEIGH_SCATTER, event.REFLECT_DIFFUSE, event.REFLECT_SPECULAR):
        m, r = ((mine.flags & bit) != 0).mean(), ((ref.flags & bit) != 0).mean()
        assert m > 0.001 and abs(m - r) < 1e-3, (hex(bit), m, r)
    print('config 4 identical fraction %.5f' % same.mean())


def test_config2_ten_million_rays_bit_exact(gpu_ready):
    geo = bench.rays_scene()
    from chroma_lite_b200.bvh import make_recursive_grid_bvh
    geo.bvh = make_recursive_grid_bvh(geo.mesh)
    n = 10000000
    o, d = bench.make_rays(geo, n)
    g = gpu.GPUGeometry(geo)
    tri, dist = gpu.intersect_mesh(g, o, d)
    tri, dist = tri.get(), dist.get()
    desc, keep = make_desc(geo)
    rg = ref_driver.RefGeometry(desc, keep)
    rtri, rdist, _ = ref_driver.intersect(rg, o, d, block=64)
    assert np.array_equal(tri, rtri), 'triangle mismatch on %d rays' % (tri != rtri).sum()
    hit = rtri >= 0
    assert np.array_equal(dist[hit].view(np.uint32), rdist[hit].view(np.uint32))
    assert 0.2 < hit.mean() < 1.0
