"""GPUPDF / GPUKernelPDF: per-channel PDF and likelihood accumulators fed by the
DAQ output (role of chroma/gpu/pdf.py:7-372 on chroma/cuda/pdf.cu), on the C ABI
(cb_pdf_*).  Same classes, methods, argument meaning and return values; the
host-side estimators follow the reference's formulas line for line, the device
side is csrc/pdf.cu.  Every accumulate call takes the GPUChannels a
GPUDaq.end_acquire() returned.
"""
import numpy as np

from .. import _lib
from .. import gpuarray as ga


def _f32(a):
    return ga.to_gpu(np.ascontiguousarray(a, dtype=np.float32))


def _per_count(values, hitcount):
    return values / np.maximum(1, hitcount)          # channels without statistics keep 0


class GPUKernelPDF(object):
    """Kernel-density estimate of the PDF value of one event's (t, q) per channel."""

    def __init__(self):
        self.lib = _lib.lib()

    # ---- first pass over the Monte Carlo: moments -> bandwidths
    def setup_moments(self, nchannels, trange, qrange, time_only=True):
        """trange / qrange: (lo, hi) of the time / charge dimension of the PDF; time_only:
        use the time observable alone (gpu/pdf.py:14-36)."""
        self.hitcount_gpu = ga.zeros(nchannels, np.uint32)
        self.tmom1_gpu, self.tmom2_gpu = ga.zeros(nchannels, np.float32), ga.zeros(nchannels, np.float32)
        self.qmom1_gpu, self.qmom2_gpu = ga.zeros(nchannels, np.float32), ga.zeros(nchannels, np.float32)
        self.trange, self.qrange = tuple(map(float, trange)), tuple(map(float, qrange))
        self.time_only = time_only

    def clear_moments(self):
        self.hitcount_gpu.fill(0)
        for a in (self.tmom1_gpu, self.tmom2_gpu, self.qmom1_gpu, self.qmom2_gpu):
            a.fill(0.0)

    def accumulate_moments(self, gpuchannels, nthreads_per_block=64):
        n = len(gpuchannels.t)
        _lib.check(self.lib.cb_pdf_accumulate_moments(
            int(self.time_only), n, gpuchannels.t.ptr, gpuchannels.q.ptr, self.trange[0], self.trange[1],
            self.qrange[0], self.qrange[1], self.hitcount_gpu.ptr, self.tmom1_gpu.ptr, self.tmom2_gpu.ptr,
            self.qmom1_gpu.ptr, self.qmom2_gpu.ptr))

    def compute_bandwidth(self, event_hit, event_time, event_charge, scale_factor=1.0):
        """Adaptive bandwidths from the accumulated moments (gpu/pdf.py:63-112): Silverman-type
        factor in d dimensions over the Gaussian density at the event's observable."""
        rho = 1.0
        n = np.maximum(self.hitcount_gpu.get(), 1)
        d = 1 if self.time_only else 2
        factor = ((4.0 / (d + 2)) / (n / scale_factor)) ** (-1.0 / (d + 4))

        def bandwidth(m1, m2, observed, clip_variance):
            mean = m1 / n
            var = m2 / n - mean ** 2
            if clip_variance:
                var = np.maximum(var, 0.0)                    # round-off can push it below zero
            rms = var ** 0.5
            with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
                density = np.minimum(1.0 / rms, (1.0 / np.sqrt(2.0 * np.pi)) * np.exp(-0.5 * ((observed - mean) / rms)) / rms)
                return factor / density * rho

        with np.errstate(divide='ignore', invalid='ignore'):
            tb = bandwidth(self.tmom1_gpu.get(), self.tmom2_gpu.get(), event_time, True)
            inv_t = np.zeros_like(tb)
            inv_t[tb > 0] = tb[tb > 0] ** -1
            self.inv_time_bandwidths_gpu = _f32(inv_t)
            if self.time_only:
                self.inv_charge_bandwidths_gpu = ga.zeros(len(inv_t), np.float32)
            else:
                qb = bandwidth(self.qmom1_gpu.get(), self.qmom2_gpu.get(), event_charge, False)
                self.inv_charge_bandwidths_gpu = _f32(qb ** -1)

    # ---- second pass: kernel evaluation at the event's observables
    def setup_kernel(self, event_hit, event_time, event_charge):
        """event_hit / event_time / event_charge: per-channel hit flag, time and charge of the
        event whose likelihood is wanted (unhit channels are ignored)."""
        self.event_hit_gpu = ga.to_gpu(np.ascontiguousarray(event_hit, dtype=np.uint32))
        self.event_time_gpu, self.event_charge_gpu = _f32(event_time), _f32(event_charge)
        self.hitcount_gpu.fill(0)
        self.time_pdf_values_gpu = ga.zeros(len(event_hit), np.float32)
        self.charge_pdf_values_gpu = ga.zeros(len(event_hit), np.float32)

    def clear_kernel(self):
        self.hitcount_gpu.fill(0)
        self.time_pdf_values_gpu.fill(0.0)
        self.charge_pdf_values_gpu.fill(0.0)

    def accumulate_kernel(self, gpuchannels, nthreads_per_block=64):
        _lib.check(self.lib.cb_pdf_accumulate_kernel_eval(
            int(self.time_only), len(self.event_hit_gpu), self.event_hit_gpu.ptr, self.event_time_gpu.ptr,
            self.event_charge_gpu.ptr, gpuchannels.t.ptr, gpuchannels.q.ptr, self.trange[0], self.trange[1],
            self.qrange[0], self.qrange[1], self.inv_time_bandwidths_gpu.ptr, self.inv_charge_bandwidths_gpu.ptr,
            self.hitcount_gpu.ptr, self.time_pdf_values_gpu.ptr, self.charge_pdf_values_gpu.ptr))

    def get_kernel_eval(self):
        """(hitcount, pdf value, uncertainty [zeros]) per channel (gpu/pdf.py:163-177)."""
        hitcount = self.hitcount_gpu.get()
        values = _per_count(self.time_pdf_values_gpu.get(), hitcount)
        if not self.time_only:
            values = values * _per_count(self.charge_pdf_values_gpu.get(), hitcount)
        return hitcount, values, np.zeros_like(values)


class GPUPDF(object):
    """(channel, t, q) histograms and the adaptive-bin PDF evaluation of one event."""

    def __init__(self):
        self.lib = _lib.lib()

    # ---- histograms
    def setup_pdf(self, nchannels, tbins, trange, qbins, qrange):
        self.events_in_histogram = 0
        self.hitcount_gpu = ga.zeros(nchannels, np.uint32)
        self.pdf_gpu = ga.zeros((nchannels, tbins, qbins), np.uint32)
        self.tbins, self.qbins = int(tbins), int(qbins)
        self.trange, self.qrange = tuple(map(float, trange)), tuple(map(float, qrange))

    def clear_pdf(self):
        self.hitcount_gpu.fill(0)
        self.pdf_gpu.fill(0)

    def add_hits_to_pdf(self, gpuchannels, nthreads_per_block=64):
        _lib.check(self.lib.cb_pdf_bin_hits(len(self.hitcount_gpu), gpuchannels.q.ptr, gpuchannels.t.ptr,
                                            self.hitcount_gpu.ptr, self.tbins, self.trange[0], self.trange[1],
                                            self.qbins, self.qrange[0], self.qrange[1], self.pdf_gpu.ptr))
        self.events_in_histogram += 1

    def get_pdfs(self):
        """1-D hit counts and the 3-D [channel, time, charge] histogram."""
        return self.hitcount_gpu.get(), self.pdf_gpu.get()

    # ---- PDF value at one event's times, bin grown until it holds min_bin_content MC hits
    def setup_pdf_eval(self, event_hit, event_time, event_charge, min_twidth, trange, min_qwidth, qrange,
                       min_bin_content=10, time_only=True):
        """The effective bin is min_twidth wide around the event's time in a channel, or as wide
        as it takes to hold min_bin_content Monte Carlo hits (gpu/pdf.py:224-287)."""
        assert time_only                                        # as in the reference: time only
        event_hit = np.asarray(event_hit)
        self.event_nhit = int(np.count_nonzero(event_hit))
        self.map_hit_offset_to_channel_id = np.flatnonzero(event_hit).astype(np.uint32)
        self.map_hit_offset_to_channel_id_gpu = ga.to_gpu(self.map_hit_offset_to_channel_id)
        self.map_channel_id_to_hit_offset = np.maximum(0, event_hit.astype(np.int64).cumsum() - 1).astype(np.uint32)
        self.event_hit_gpu = ga.to_gpu(np.ascontiguousarray(event_hit, dtype=np.uint32))
        self.event_time_gpu, self.event_charge_gpu = _f32(event_time), _f32(event_charge)
        self.eval_hitcount_gpu = ga.zeros(len(event_hit), np.uint32)
        self.eval_bincount_gpu = ga.zeros(len(event_hit), np.uint32)
        self.nearest_mc_gpu = ga.empty(max(self.event_nhit * min_bin_content, 1), np.float32)
        self.nearest_mc_gpu.fill(1e9)
        self.min_twidth, self.min_qwidth = float(min_twidth), min_qwidth
        self.trange, self.qrange = tuple(map(float, trange)), qrange
        self.min_bin_content, self.time_only = min_bin_content, time_only

    def clear_pdf_eval(self):
        self.eval_hitcount_gpu.fill(0)
        self.eval_bincount_gpu.fill(0)
        self.nearest_mc_gpu.fill(1e9)

    def accumulate_pdf_eval(self, gpuchannels, nthreads_per_block=64, max_blocks=10000):
        """One launch (the reference: work-queue fill, accumulate_bincount, sync,
        accumulate_nearest_neighbor_block, sync)."""
        _lib.check(self.lib.cb_pdf_accumulate_eval(
            len(self.event_hit_gpu), int(gpuchannels.ndaq), self.event_nhit, self.event_hit_gpu.ptr,
            self.event_time_gpu.ptr, gpuchannels.t.ptr, self.eval_hitcount_gpu.ptr, self.eval_bincount_gpu.ptr,
            self.min_twidth, self.trange[0], self.trange[1], int(self.min_bin_content),
            self.map_hit_offset_to_channel_id_gpu.ptr, self.nearest_mc_gpu.ptr))

    def get_pdf_eval(self):
        """(hitcount, pdf value, uncertainty) per channel (gpu/pdf.py:332-372): bins with
        min_bin_content hits inside min_twidth use the count, the others the distance to the
        farthest of the nearest hits collected so far."""
        evhit = self.event_hit_gpu.get().astype(bool)
        hitcount, bincount = self.eval_hitcount_gpu.get(), self.eval_bincount_gpu.get()
        m = self.min_bin_content
        value = np.zeros(len(hitcount), dtype=float)
        frac = np.zeros_like(value)
        high = bincount >= m
        value[high] = bincount[high].astype(float) / hitcount[high] / self.min_twidth
        frac[high] = 1.0 / np.sqrt(bincount[high])
        low = ~high & (hitcount > 0) & evhit
        nearest = np.full((len(hitcount), m), 1e9, dtype=np.float32)
        nearest[self.map_hit_offset_to_channel_id] = self.nearest_mc_gpu.get()[:self.event_nhit * m].reshape(self.event_nhit, m)
        last = np.maximum(0, (nearest < 1e9).sum(axis=1) - 1)          # clamped: a channel may have none yet
        distance = nearest[np.arange(len(last)), last]
        value[low] = (last[low] + 1).astype(float) / hitcount[low] / distance[low] / 2.0
        frac[low] = 1.0 / np.sqrt(last[low] + 1)
        return hitcount, value, value * frac
