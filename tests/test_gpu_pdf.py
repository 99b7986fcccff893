"""PDF / likelihood accumulators: engine (csrc/pdf.cu through gpu.GPUPDF /
gpu.GPUKernelPDF) vs the reference's own pdf.cu kernels (oracle/_ref/pdf.cubin) on the same
DAQ outputs -- integer state bit-exact, float state bit-exact or within 1e-6 relative -- and vs
the NumPy restatement.  The channels come from real acquisitions of the tiny detector plus
synthetic ones that reach every branch (range edges, unhit channels, ndaq > 1, list overflow)."""
import numpy as np
import pytest

from chroma_lite_b200 import gpu
from chroma_lite_b200 import gpuarray as ga
from oracle import ref_driver, pdf_oracle as po
import scenes

pytestmark = pytest.mark.gpu


def synthetic(n, ndaq, seed, hit_fraction=0.6):
    rng = np.random.default_rng(seed)
    t = rng.normal(40.0, 6.0, (ndaq, n)).astype(np.float32)
    q = (rng.gamma(4.0, 1.5, (ndaq, n)) * rng.choice([1.0, 0.999], size=(ndaq, n))).astype(np.float32)
    miss = rng.uniform(size=(ndaq, n)) > hit_fraction
    t[miss] = 1e9
    q[miss] = 0
    # values sitting exactly on the range edges
    t[0, :8] = [20.0, 60.0, 19.999999, 59.999996, 40.0, 40.0, 1e8, 99999992.0]
    q[0, :8] = [0.0, 14.0, 13.999999, 14.000001, -1.0, 3.5, 2.0, 2.0]
    return t.reshape(-1), q.reshape(-1)


def channels_of(t, q, ndaq=1):
    return gpu.GPUChannels(ga.to_gpu(t), ga.to_gpu(q), ga.to_gpu(np.zeros(len(t), np.uint32)), ndaq=ndaq,
                           stride=len(t) // ndaq)


def daq_acquisitions(gpu_ready, count):
    """real (t, q) per channel: `count` events in the tiny detector"""
    geo = scenes.tiny_detector()
    g = gpu.GPUDetector(geo)
    daq = gpu.GPUDaq(g)
    out = []
    for k in range(count):
        ph = scenes.point_source(60000, seed=200 + k, wl_range=(300, 600))
        rng = gpu.get_rng_states(len(ph), seed=300 + k)
        gp = gpu.GPUPhotons(ph)
        gp.propagate(g, rng, nthreads_per_block=256, max_blocks=(len(ph) + 255) // 256, max_steps=50)
        daq.begin_acquire()
        daq.acquire(gp, rng, nthreads_per_block=256, max_blocks=(len(ph) + 255) // 256)
        ch = daq.end_acquire()
        out.append((ch.t.get().copy(), ch.q.get().copy()))
    return g.nchannels, out


def close(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    exact = (a.view(np.uint32) == b.view(np.uint32)).mean()
    assert np.allclose(a, b, rtol=1e-6, atol=1e-12), '%s: max rel diff %.3g' % (what, np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30)))
    return exact


def test_histograms_vs_reference_kernel(gpu_ready):
    nch, acqs = daq_acquisitions(gpu_ready, 3)
    trange, qrange, tb, qb = (0.0, 200.0), (0.0, 10.0), 50, 10
    mine, ref = gpu.GPUPDF(), ref_driver.RefPDF()
    mine.setup_pdf(nch, tb, trange, qb, qrange)
    ref.setup_pdf(nch, tb, trange, qb, qrange)
    o_hc, o_pdf = np.zeros(nch, np.uint32), np.zeros((nch, tb, qb), np.uint32)
    for t, q in acqs:
        mine.add_hits_to_pdf(channels_of(t, q))
        ref.add_hits_to_pdf(t, q)
        po.bin_hits(q, t, o_hc, o_pdf, trange, qrange)
    hc, pdf = mine.get_pdfs()
    rhc, rpdf = ref.get_pdfs()
    assert np.array_equal(hc, rhc) and np.array_equal(pdf, rpdf)
    assert hc.sum() > 30 and mine.events_in_histogram == 3
    assert np.array_equal(hc, o_hc) and np.array_equal(pdf, o_pdf)
    # synthetic channels incl. the range edges, more bins
    n = 3000
    mine.setup_pdf(n, 40, (20.0, 60.0), 14, (0.0, 14.0))
    ref.setup_pdf(n, 40, (20.0, 60.0), 14, (0.0, 14.0))
    for k in range(4):
        t, q = synthetic(n, 1, k)
        mine.add_hits_to_pdf(channels_of(t, q))
        ref.add_hits_to_pdf(t, q)
    assert all(np.array_equal(a, b) for a, b in zip(mine.get_pdfs(), ref.get_pdfs()))
    mine.clear_pdf()
    assert mine.get_pdfs()[1].sum() == 0


@pytest.mark.parametrize('time_only', [True, False])
def test_kernel_pdf_vs_reference_kernel(gpu_ready, time_only):
    n = 5000
    trange, qrange = (20.0, 60.0), (0.0, 14.0)
    mine = gpu.GPUKernelPDF()
    mine.setup_moments(n, trange, qrange, time_only=time_only)
    ref = ref_driver.RefKernelPDF(n, trange, qrange, time_only=time_only)
    o = [np.zeros(n, np.uint32)] + [np.zeros(n, np.float32) for _ in range(4)]
    for k in range(6):
        t, q = synthetic(n, 1, 10 + k)
        mine.accumulate_moments(channels_of(t, q))
        ref.accumulate_moments(t, q)
        po.accumulate_moments(time_only, t, q, trange, qrange, *o)
    got = (mine.hitcount_gpu.get(), mine.tmom1_gpu.get(), mine.tmom2_gpu.get(), mine.qmom1_gpu.get(), mine.qmom2_gpu.get())
    want = ref.moments()
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[0], o[0])
    for a, b, c, name in zip(got[1:], want[1:], o[1:], ('tmom1', 'tmom2', 'qmom1', 'qmom2')):
        assert close(a, b, name) > 0.99
        assert np.allclose(a, c, rtol=1e-5)
    # kernel evaluation with the bandwidths the host estimator derives from those moments
    rng = np.random.default_rng(3)
    event_hit = rng.uniform(size=n) < 0.7
    event_time = rng.normal(40.0, 5.0, n).astype(np.float32)
    event_charge = rng.gamma(4.0, 1.5, n).astype(np.float32)
    mine.compute_bandwidth(event_hit, event_time, event_charge)
    inv_t, inv_q = mine.inv_time_bandwidths_gpu.get(), mine.inv_charge_bandwidths_gpu.get()
    inv_t[:50] = 0.0                                        # zero bandwidth: flat normalisation branch
    mine.inv_time_bandwidths_gpu = ga.to_gpu(inv_t)
    assert np.isfinite(inv_t).all() and (inv_t > 0).mean() > 0.5
    mine.setup_kernel(event_hit, event_time, event_charge)
    ref.setup_kernel(event_hit, event_time, event_charge, inv_t, inv_q)
    o_hc, o_t, o_q = np.zeros(n, np.uint32), np.zeros(n, np.float32), np.zeros(n, np.float32)
    for k in range(5):
        t, q = synthetic(n, 1, 40 + k)
        mine.accumulate_kernel(channels_of(t, q))
        ref.accumulate_kernel(t, q)
        po.accumulate_kernel_eval(time_only, event_hit, event_time, event_charge, t, q, trange, qrange,
                                  inv_t, np.nan_to_num(inv_q, posinf=0.0), o_hc, o_t, o_q)
    rhc, rt, rq = ref.kernel_state()
    assert np.array_equal(mine.hitcount_gpu.get(), rhc) and np.array_equal(rhc, o_hc)
    fin = np.isfinite(rt)
    assert fin.mean() > 0.95
    exact = close(mine.time_pdf_values_gpu.get()[fin], rt[fin], 'time pdf values')
    assert exact > 0.95, exact
    assert np.allclose(mine.time_pdf_values_gpu.get()[fin], o_t[fin], rtol=2e-4, atol=1e-7)
    if not time_only:
        finq = np.isfinite(rq)
        assert close(mine.charge_pdf_values_gpu.get()[finq], rq[finq], 'charge pdf values') > 0.95
    hc, val, err = mine.get_kernel_eval()
    assert (val[event_hit & (hc > 0)] > 0).mean() > 0.9 and not err.any()


@pytest.mark.parametrize('ndaq,m', [(1, 10), (8, 6), (40, 25), (100, 3)])
def test_pdf_eval_vs_reference_kernels(gpu_ready, ndaq, m):
    n = 2000
    rng = np.random.default_rng(ndaq)
    event_hit = rng.uniform(size=n) < 0.4
    event_hit[:3] = [True, False, True]
    event_time = rng.normal(40.0, 4.0, n).astype(np.float32)
    trange, width = (20.0, 60.0), 1.0
    mine, ref = gpu.GPUPDF(), ref_driver.RefPDF()
    mine.setup_pdf_eval(event_hit, event_time, event_time, width, trange, 1.0, (0.0, 14.0), min_bin_content=m)
    ref.setup_pdf_eval(event_hit, event_time, width, trange, min_bin_content=m)
    nhit = int(event_hit.sum())
    o_hc, o_bc, o_near = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.full((nhit, m), 1e9, np.float32)
    for k in range(4):
        t, q = synthetic(n, ndaq, 70 + k, hit_fraction=0.5)
        mine.accumulate_pdf_eval(channels_of(t, q, ndaq))
        ref.accumulate_pdf_eval(t, ndaq)
        po.accumulate_pdf_eval(event_hit, event_time, t, ndaq, o_hc, o_bc, o_near, width, trange, m)
        rhc, rbc, rnear = ref.eval_state()
        assert np.array_equal(mine.eval_hitcount_gpu.get(), rhc), 'hitcount after acquisition %d' % k
        assert np.array_equal(mine.eval_bincount_gpu.get(), rbc), 'bincount after acquisition %d' % k
        assert np.array_equal(mine.nearest_mc_gpu.get().view(np.uint32), rnear.view(np.uint32)), 'nearest list %d' % k
    assert np.array_equal(rhc, o_hc) and np.array_equal(rbc, o_bc)
    assert np.array_equal(rnear.reshape(nhit, m).view(np.uint32), o_near.view(np.uint32))
    hc, val, err = mine.get_pdf_eval()
    assert (val[event_hit & (hc > 0)] > 0).all() and (val[~event_hit] == 0).all()
    mine.clear_pdf_eval()
    assert mine.eval_hitcount_gpu.get().sum() == 0 and (mine.nearest_mc_gpu.get() > 1e8).all()


def test_pdf_errors_are_loud(gpu_ready):
    from chroma_lite_b200 import _lib
    p = gpu.GPUPDF()
    n = 64
    hit = np.ones(n, bool)
    p.setup_pdf_eval(hit, np.zeros(n, np.float32), np.zeros(n, np.float32), 1.0, (0.0, 1.0), 1.0, (0.0, 1.0),
                     min_bin_content=200000)
    t = np.zeros(n, np.float32)
    with pytest.raises(_lib.ChromaB200Error):
        p.accumulate_pdf_eval(channels_of(t, t))


def test_simulation_create_and_eval_pdf(gpu_ready):
    """The upstream Simulation.create_pdf / eval_pdf / eval_kernel front end on the tiny detector."""
    from chroma_lite_b200 import sim
    det = scenes.tiny_detector()
    s = sim.Simulation(det, seed=5, nthreads_per_block=256, max_blocks=512)
    mc = [scenes.point_source(40000, seed=400 + k, wl_range=(300, 600)) for k in range(3)]
    hitcount, pdf = s.create_pdf(mc, 40, (0.0, 200.0), 10, (0.0, 10.0), nreps=2)
    assert pdf.shape == (s.gpu_geometry.nchannels, 40, 10)
    assert np.array_equal(pdf.sum(axis=(1, 2)), hitcount) and hitcount.sum() > 50
    assert s.gpu_pdf.events_in_histogram == 6
    # the event of interest: one more simulated event's channels
    ev = next(s.simulate([scenes.point_source(40000, seed=999, wl_range=(300, 600))], run_daq=True, max_steps=100,
                         keep_flat_hits=False, keep_hits=False))
    hc, val, err = s.eval_pdf(ev.channels, mc, 2.0, (0.0, 200.0), 1.0, (0.0, 10.0), min_bin_content=5, ndaq=2)
    seen = ev.channels.hit & (hc > 0)
    assert seen.sum() > 5 and (val[seen] > 0).all() and np.isfinite(val).all() and (err[seen] > 0).all()
    assert (val[~ev.channels.hit] == 0).all()
    hc2, val2, _ = s.eval_kernel(ev.channels, mc, (0.0, 200.0), (0.0, 10.0), time_only=True)
    ok = ev.channels.hit & (hc2 > 1)
    assert ok.sum() > 5 and np.isfinite(val2[ok]).all() and (val2[ok] >= 0).all()
