"""PDF accumulators, CPU side: the NumPy restatement of pdf.cu behaves like the textbook
definitions (histogram, moments, k-nearest list), and the host estimators of gpu/pdf.py
(bandwidths, PDF values) are evaluated on arrays without a GPU."""
import numpy as np

from oracle import pdf_oracle as po


def fake_channels(n, ndaq, seed, hit_fraction=0.6):
    rng = np.random.default_rng(seed)
    t = rng.normal(40.0, 6.0, (ndaq, n)).astype(np.float32)
    q = np.round(rng.gamma(4.0, 1.5, (ndaq, n))).astype(np.float32)
    miss = rng.uniform(size=(ndaq, n)) > hit_fraction
    t[miss] = 1e9
    q[miss] = 0
    return t.reshape(-1), q.reshape(-1)


def test_bin_hits_is_a_histogram():
    n, tb, qb = 500, 12, 7
    hitcount = np.zeros(n, np.uint32)
    pdf = np.zeros((n, tb, qb), np.uint32)
    total = 0
    for k in range(5):
        t, q = fake_channels(n, 1, k)
        po.bin_hits(q, t, hitcount, pdf, (20.0, 60.0), (0.0, 14.0))
        ok = (t >= 20) & (t < 60) & (np.trunc(q) >= 0) & (np.trunc(q) < 14)
        total += ok.sum()
    assert pdf.sum() == total == hitcount.sum()
    assert np.array_equal(pdf.sum(axis=(1, 2)), hitcount)
    # one channel against numpy's histogram2d
    t, q = fake_channels(n, 1, 99)
    h2 = np.zeros((n, tb, qb), np.uint32)
    po.bin_hits(q, t, np.zeros(n, np.uint32), h2, (20.0, 60.0), (0.0, 14.0))
    ch = int(np.flatnonzero(h2.sum(axis=(1, 2)))[0])
    ref, _, _ = np.histogram2d([t[ch]], [np.trunc(q[ch])], bins=(tb, qb), range=((20, 60), (0, 14)))
    assert np.array_equal(h2[ch], ref.astype(np.uint32))


def test_moments_and_kernel_eval():
    n = 400
    mom0 = np.zeros(n, np.uint32)
    t1, t2, q1, q2 = (np.zeros(n, np.float32) for _ in range(4))
    ts = []
    for k in range(40):
        t, q = fake_channels(n, 1, 100 + k)
        po.accumulate_moments(False, t, q, (0.0, 100.0), (0.0, 50.0), mom0, t1, t2, q1, q2)
        ts.append(np.where(t < 1e8, t, np.nan))
    ts = np.array(ts)
    assert np.array_equal(mom0, (~np.isnan(ts)).sum(axis=0))
    assert np.allclose(t1 / np.maximum(mom0, 1), np.nanmean(ts, axis=0), rtol=1e-4)
    # kernel estimate of a Gaussian time PDF at its mean: ~ 1/(sigma sqrt(2 pi)) with a narrow kernel
    hit = np.ones(n, np.uint32)
    ev_t = np.full(n, 40.0, np.float32)
    inv_t = np.full(n, 1.0 / 1.5, np.float32)
    hc = np.zeros(n, np.uint32)
    tv, qv = np.zeros(n, np.float32), np.zeros(n, np.float32)
    for k in range(300):
        t, q = fake_channels(n, 1, 1000 + k, hit_fraction=1.0)
        po.accumulate_kernel_eval(True, hit, ev_t, ev_t, t, q, (0.0, 100.0), (0.0, 50.0), inv_t, inv_t, hc, tv, qv)
    est = (tv / np.maximum(hc, 1)).mean()
    expect = 1.0 / np.sqrt(2 * np.pi * (6.0 ** 2 + 1.5 ** 2))
    assert abs(est - expect) / expect < 0.03


def test_pdf_eval_keeps_the_nearest_distances():
    n, ndaq, m = 60, 16, 6
    rng = np.random.default_rng(5)
    event_hit = rng.uniform(size=n) < 0.5
    event_time = rng.normal(40.0, 3.0, n).astype(np.float32)
    nhit = int(event_hit.sum())
    hitcount, bincount = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    nearest = np.full((nhit, m), 1e9, np.float32)
    everything = [[] for _ in range(n)]
    for k in range(4):
        t, _ = fake_channels(n, ndaq, 50 + k)
        po.accumulate_pdf_eval(event_hit, event_time, t, ndaq, hitcount, bincount, nearest, 1.0, (0.0, 100.0), m)
        tt = t.reshape(ndaq, n)
        for ch in range(n):
            everything[ch] += [abs(np.float32(x - event_time[ch])) for x in tt[:, ch] if x < 1e8]
    rows = np.flatnonzero(event_hit)
    for r, ch in enumerate(rows):
        assert hitcount[ch] == len(everything[ch])
        assert bincount[ch] == sum(d < 0.5 for d in everything[ch])
        got = nearest[r][nearest[r] < 1e8]
        assert np.all(np.diff(got) >= 0)
        if bincount[ch] < m:            # the list never stopped filling: it holds the m smallest distances
            assert np.allclose(got, np.sort(np.array(everything[ch], np.float32))[:m])
    assert (hitcount[~event_hit] > 0).any() and (bincount[~event_hit] == 0).all()


def test_host_estimators_without_gpu(monkeypatch):
    """get_pdf_eval / get_kernel_eval / compute_bandwidth are host arithmetic: drive them with
    stand-in device arrays."""
    from chroma_lite_b200.gpu import pdf as gpdf

    class Arr(object):
        def __init__(self, a):
            self.a = np.asarray(a)

        def get(self):
            return self.a.copy()

        def __len__(self):
            return len(self.a)

    monkeypatch.setattr(gpdf._lib, 'lib', lambda: None)
    monkeypatch.setattr(gpdf.ga, 'to_gpu', lambda a: Arr(a))
    monkeypatch.setattr(gpdf.ga, 'zeros', lambda n, dt: Arr(np.zeros(n, dt)))
    p = gpdf.GPUPDF()
    p.min_bin_content, p.min_twidth, p.time_only, p.event_nhit = 4, 2.0, True, 2
    p.event_hit_gpu = Arr([1, 0, 1])
    p.eval_hitcount_gpu = Arr(np.array([100, 50, 10], np.uint32))
    p.eval_bincount_gpu = Arr(np.array([8, 0, 1], np.uint32))
    p.map_hit_offset_to_channel_id = np.array([0, 2], np.uint32)
    p.nearest_mc_gpu = Arr(np.array([0.1, 0.2, 0.3, 0.4, 0.5, 3.0, 1e9, 1e9], np.float32))
    hc, val, err = p.get_pdf_eval()
    assert np.isclose(val[0], 8 / 100 / 2.0) and np.isclose(err[0], val[0] / np.sqrt(8))
    assert val[1] == 0
    assert np.isclose(val[2], 2 / 10 / 3.0 / 2.0) and np.isclose(err[2], val[2] / np.sqrt(2))
    k = gpdf.GPUKernelPDF()
    k.time_only = False
    k.hitcount_gpu = Arr(np.array([4, 0], np.uint32))
    k.time_pdf_values_gpu, k.charge_pdf_values_gpu = Arr(np.array([2.0, 0.0], np.float32)), Arr(np.array([1.0, 0.0], np.float32))
    hc, val, err = k.get_kernel_eval()
    assert np.allclose(val, [0.5 * 0.25, 0.0]) and not err.any()
    k.tmom1_gpu, k.tmom2_gpu = Arr(np.array([160.0, 0.0], np.float32)), Arr(np.array([6500.0, 0.0], np.float32))
    k.qmom1_gpu, k.qmom2_gpu = Arr(np.array([20.0, 0.0], np.float32)), Arr(np.array([104.0, 0.0], np.float32))
    k.compute_bandwidth(np.array([1, 0]), np.array([41.0, 0.0]), np.array([5.0, 0.0]))
    inv_t = k.inv_time_bandwidths_gpu.get()
    assert inv_t[0] > 0 and np.isfinite(inv_t[0])


def test_likelihood_host_layer_without_gpu():
    """chroma/likelihood.py:47-180 over a stand-in simulation: hit / not-hit terms with the half-count floor,
    flat-density floor for channels without Monte Carlo data, mean and standard error of the kernel estimates."""
    from chroma_lite_b200 import event
    from chroma_lite_b200.likelihood import Likelihood
    hit = np.array([True, True, False, False, True])
    ev = event.Event(channels=event.Channels(hit, np.array([10.0, 20.0, 1e9, 1e9, 30.0], np.float32),
                                             np.array([1.0, 1.0, 0.0, 0.0, 2.0], np.float32)))

    class FakeSim(object):
        def __init__(self):
            self.calls = []

        def eval_pdf(self, channels, it, min_twidth, trange, min_qwidth, qrange, **kw):
            self.calls.append(('pdf', len(list(it)), min_twidth, kw))
            return (np.array([80, 0, 8, 0, 40]), np.array([0.05, 0.0, 0.0, 0.0, np.nan], np.float32),
                    np.array([0.01, 0.0, 0.0, 0.0, 0.0], np.float32))

        def eval_kernel(self, channels, events, trange, qrange, **kw):
            self.calls.append(('kernel', len(events), kw))
            k = len([c for c in self.calls if c[0] == 'kernel'])
            return np.zeros(5), np.array([0.01 * k, 0.02, 0.0, 0.0, 0.04], np.float32), np.zeros(5, np.float32)
    fs = FakeSim()
    lk = Likelihood(fs, ev, trange=(0.0, 100.0))
    nll = lk.eval(iter(range(1000)), nevals=4, nreps=2, ndaq=10)                      # ntotal = 80
    assert fs.calls[0][:3] == ('pdf', 4, 0.2) and fs.calls[0][3]['min_bin_content'] == 320
    p = np.array([80 / 80, 0.5 / 80, 1 - 8 / 80, 1.0, 40 / 80])                        # floor for the hit channel with no MC hits
    want = -(np.log(p).sum() + np.log([0.05, 0.01, 0.01]).sum())                       # 1/(100-0) where the PDF is 0 / NaN
    assert np.isclose(float(nll), want, rtol=1e-6) and lk.channels_without_data == 2
    gen = iter(range(1000))
    res = lk.eval_kernel(gen, nevals=3, nreps=1, ndaq=1, navg=4)
    lls = [np.log([0.01 * k, 0.02, 0.04]).sum() for k in (1, 2, 3, 4)]
    assert np.isclose(res.nominal_value, -np.mean(lls)) and np.isclose(res.std_dev, np.std(lls) / 2.0)
    assert next(gen) == 12 and [c[1] for c in fs.calls if c[0] == 'kernel'] == [3, 3, 3, 3]
    lk2 = Likelihood(fs, ev, trange=(0.0, 100.0), qrange=(0.0, 10.0), time_only=False)
    assert np.isclose(lk2._pdf_floor(), 1e-3)
