"""CPU restatement (NumPy, float32 where the kernels use float) of the reference's PDF
accumulators, chroma/cuda/pdf.cu.  TEST INFRASTRUCTURE ONLY: imported by tests/ only.

Pinning: the reference's own test for this code (test/test_pdf.py) needs the removed
Geant4 generator and holds no vectors, so this restatement is pinned against the
reference kernels run on the GPU (oracle/_ref/pdf.cubin, tests/test_gpu_pdf.py) --
"parity unpinned" by reference fixtures, pinned by the reference itself.

Every function takes and returns host arrays and ADDS one acquisition, like the kernels.
"""
import numpy as np
from scipy.special import erf

F = np.float32


def bin_hits(q, t, hitcount, pdf, trange, qrange):
    """bin_hits (pdf.cu:9-32): the charge is truncated to an unsigned integer first; row-major
    (channel, tbin, qbin) histogram."""
    nch, tbins, qbins = pdf.shape
    q = np.asarray(q, F)[:nch]
    t = np.asarray(t, F)[:nch]
    qi = np.clip(np.trunc(q), 0, 2 ** 32 - 1).astype(np.uint32).astype(F)     # cvt.rzi.u32.f32 saturates
    tmin, tmax, qmin, qmax = F(trange[0]), F(trange[1]), F(qrange[0]), F(qrange[1])
    ok = (t < 1e8) & (t >= tmin) & (t < tmax) & (qi >= qmin) & (qi < qmax)
    tb = ((t - tmin) / (tmax - tmin) * F(tbins)).astype(np.int64)
    qb = ((qi - qmin) / (qmax - qmin) * F(qbins)).astype(np.int64)
    ch = np.flatnonzero(ok)
    hitcount[ch] += 1
    np.add.at(pdf, (ch, tb[ch], qb[ch]), 1)
    return tb, qb, ok


def accumulate_moments(time_only, t, q, trange, qrange, mom0, t1, t2, q1, q2):
    """accumulate_moments (pdf.cu:223-266)."""
    t, q = np.asarray(t, F), np.asarray(q, F)
    ok = ~((t < F(trange[0])) | (t > F(trange[1])))
    if not time_only:
        ok &= ~((q < F(qrange[0])) | (q > F(qrange[1])))
    mom0[ok] += 1
    t1[ok] += t[ok]
    t2[ok] += t[ok] * t[ok]
    if not time_only:
        q1[ok] += q[ok]
        q2[ok] += q[ok] * q[ok]


def _window_norm(lo, hi, mc, inv_bw):
    norm = np.full(mc.shape, F(hi) - F(lo), dtype=F)
    pos = inv_bw > 0
    s = F(0.70710678118654746)
    lo_arg = (F(lo) - mc) * inv_bw * s
    hi_arg = (F(hi) - mc) * inv_bw * s
    norm[pos] = ((erf(hi_arg[pos].astype(np.float64)) - erf(lo_arg[pos].astype(np.float64))) * 1.2533141373155001).astype(F)
    return norm


def accumulate_kernel_eval(time_only, event_hit, event_time, event_charge, t, q, trange, qrange, inv_t, inv_q,
                           hitcount, tval, qval):
    """accumulate_kernel_eval (pdf.cu:271-368)."""
    t, q = np.asarray(t, F), np.asarray(q, F)
    ok = ~((t < F(trange[0])) | (t > F(trange[1])))
    if not time_only:
        ok &= ~((q < F(qrange[0])) | (q > F(qrange[1])))
    hitcount[ok] += 1
    ev = ok & (np.asarray(event_hit) != 0)
    arg = (t - np.asarray(event_time, F)) * inv_t
    term = np.exp(F(-0.5) * arg * arg)
    if time_only:
        term = term * inv_t
    with np.errstate(divide='ignore', invalid='ignore'):
        tval[ev] += (term / _window_norm(trange[0], trange[1], t, inv_t))[ev]
        if not time_only:
            arg = (q - np.asarray(event_charge, F)) * inv_q
            qval[ev] += (np.exp(F(-0.5) * arg * arg) / _window_norm(qrange[0], qrange[1], q, inv_q))[ev]


def accumulate_pdf_eval(event_hit, event_time, mc_time, ndaq, hitcount, bincount, nearest, min_twidth, trange, m):
    """accumulate_bincount (pdf.cu:34-96) followed by accumulate_nearest_neighbor(_block)
    (pdf.cu:98-219).  mc_time: [ndaq * nchannels]; nearest: [nhit, m] sorted rows, 1e9 = unused."""
    event_hit = np.asarray(event_hit)
    nch = len(event_hit)
    mc = np.asarray(mc_time, F).reshape(ndaq, nch)
    ev_t = np.asarray(event_time, F)
    hit_rows = np.flatnonzero(event_hit)
    row_of = {int(c): r for r, c in enumerate(hit_rows)}
    for ch in range(nch):
        queued = []
        for i in range(ndaq):
            tt = mc[i, ch]
            if tt >= 1e8 or tt < F(trange[0]) or tt > F(trange[1]):
                continue
            hitcount[ch] += 1
            if not event_hit[ch]:
                continue
            dist = F(abs(F(tt - ev_t[ch])))
            if dist < F(min_twidth) / F(2.0):
                bincount[ch] += 1
            if bincount[ch] < m:
                queued.append(dist)
        if event_hit[ch] and queued:
            row = nearest[row_of[ch]]
            valid = []
            for d in row:                       # up to the first unused slot
                if d > 1e8:
                    break
                valid.append(d)
            merged = np.sort(np.array(valid + queued, dtype=F), kind='stable')[:m]
            row[:len(merged)] = merged
