"""Negative log likelihood of a detector event given a photon-event generator: the host layer above the
PDF accumulators (role of chroma/likelihood.py:7-180).

The generators here yield photon events (event.Event / event.Photons): the reference's yield Geant4
vertices, and its generator package is not part of this fork.  Results carry an uncertainty like the
reference's ``ufloat`` (the `uncertainties` package is used when it is installed, a two-field stand-in
otherwise).
"""
from itertools import islice
from math import sqrt

import numpy as np


class ValueWithUncertainty(object):
    """Stand-in for uncertainties.ufloat: nominal_value, std_dev, unary minus, float()."""

    def __init__(self, nominal_value, std_dev=0.0):
        self.nominal_value, self.std_dev = float(nominal_value), float(std_dev)

    def __neg__(self):
        return ValueWithUncertainty(-self.nominal_value, self.std_dev)

    def __float__(self):
        return self.nominal_value

    def __repr__(self):
        return '%g+/-%g' % (self.nominal_value, self.std_dev)


def ufloat(value, std_dev=0.0):
    try:
        from uncertainties import ufloat as _ufloat
        return _ufloat(value, std_dev)
    except ImportError:
        return ValueWithUncertainty(value, std_dev)


class Likelihood(object):
    "Evaluate likelihoods for detector events (chroma/likelihood.py:7-45)."

    def __init__(self, sim, event=None, tbins=100, trange=(-0.5, 999.5), qbins=10, qrange=(-0.5, 49.5),
                 time_only=True):
        self.sim = sim
        self.tbins, self.trange, self.qbins, self.qrange, self.time_only = tbins, trange, qbins, qrange, time_only
        if event is not None:
            self.set_event(event)

    def set_event(self, event):
        "Set the detector event being reconstructed."
        self.event = event

    def _pdf_floor(self):
        floor = 1.0 / (self.trange[1] - self.trange[0])
        return floor if self.time_only else floor / (self.qrange[1] - self.qrange[0])

    def _floor_bad_values(self, pdf_prob, pdf_prob_uncert):
        """Zero / NaN densities become the flat density over the range; returns how many HIT channels had none."""
        bad = (pdf_prob <= 0.0) | np.isnan(pdf_prob)
        pdf_prob[bad] = self._pdf_floor()
        pdf_prob_uncert[bad] = self._pdf_floor()
        return int((bad & self.event.channels.hit).sum())

    def eval_channel_vbin(self, vertex_generator, nevals, nreps=16, ndaq=50):
        """(hit probability, PDF value, PDF uncertainty) per channel with the variable-bin method
        (chroma/likelihood.py:47-86): 0.2 ns minimum bin, at least 320 Monte Carlo entries per bin."""
        ntotal = nevals * nreps * ndaq
        hitcount, pdf_prob, pdf_prob_uncert = self.sim.eval_pdf(
            self.event.channels, islice(vertex_generator, nevals), 0.2, self.trange, 1, self.qrange,
            nreps=nreps, ndaq=ndaq, time_only=self.time_only, min_bin_content=320)
        hit_prob = hitcount.astype(np.float32) / ntotal
        pdf_prob = np.array(pdf_prob, dtype=np.float64)
        pdf_prob_uncert = np.array(pdf_prob_uncert, dtype=np.float64)
        self.channels_without_data = self._floor_bad_values(pdf_prob, pdf_prob_uncert)
        return hit_prob, pdf_prob, pdf_prob_uncert

    def eval(self, vertex_generator, nevals, nreps=16, ndaq=50):
        """-log L of the event for the source `vertex_generator` describes (chroma/likelihood.py:88-117):
        hit / not-hit probabilities of all channels (floored at half a count) plus the log densities of
        the observed times (and charges) of the hit channels."""
        ntotal = nevals * nreps * ndaq
        hit_prob, pdf_prob, _ = self.eval_channel_vbin(vertex_generator, nevals, nreps, ndaq)
        hit = np.asarray(self.event.channels.hit, dtype=bool)
        hit_prob = hit_prob.astype(np.float64)
        hit_prob[~hit] = 1.0 - hit_prob[~hit]
        hit_prob = np.maximum(hit_prob, 0.5 / ntotal)
        log_likelihood = np.log(hit_prob).sum() + np.log(pdf_prob[hit]).sum()
        return -ufloat(log_likelihood, 0.0)

    def eval_kernel(self, vertex_generator, nevals, nreps=16, ndaq=50, navg=10, oversample_factor=1.0):
        """Kernel-density version (chroma/likelihood.py:119-180): `navg` independent estimates of the
        density term from `nevals` events each; mean and standard error of the mean.  Like the
        reference, the hit / not-hit term is left out here.  The bandwidths are set per estimate from
        the same events (Simulation.eval_kernel runs the moment pass first)."""
        hit = np.asarray(self.event.channels.hit, dtype=bool)
        mom0, mom1, mom2 = 0, 0.0, 0.0
        for _ in range(navg):
            events = list(islice(vertex_generator, nevals))
            if not events:
                break
            _, pdf_prob, pdf_prob_uncert = self.sim.eval_kernel(
                self.event.channels, events, self.trange, self.qrange, nreps=nreps, ndaq=ndaq,
                time_only=self.time_only, scale_factor=oversample_factor)
            pdf_prob = np.array(pdf_prob, dtype=np.float64)
            pdf_prob_uncert = np.array(pdf_prob_uncert, dtype=np.float64)
            self.channels_without_data = self._floor_bad_values(pdf_prob, pdf_prob_uncert)
            log_likelihood = np.log(pdf_prob[hit]).sum()
            if np.isfinite(log_likelihood):
                mom0 += 1
                mom1 += log_likelihood
                mom2 += log_likelihood ** 2
        if mom0 == 0:
            raise ValueError('no finite likelihood estimate')
        avg = mom1 / mom0
        rms = max(mom2 / mom0 - avg ** 2, 0.0) ** 0.5
        return ufloat(-avg, rms / sqrt(mom0))
