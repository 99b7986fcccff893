"""Detector: a Geometry some of whose solids are read-out channels.

Array contract of chroma/detector.py:5-141 (what GPUDetector and GPUDaq consume):
  solid_id_to_channel_index[solid]   channel of a solid, -1 for passive solids
  channel_index_to_solid_id / _channel_type / _position   per channel
  time_cdf, charge_cdf               (x, y) pairs sampled by the DAQ kernel (chroma/cuda/daq.cu:55-66)
The channel bookkeeping and the response curves are this package's own code; only the attribute and
method names are the reference's, so that the reference's Detector objects and these are
interchangeable as inputs.
"""
import numpy as np

from .geometry import Geometry

PASSIVE = -1


def cdf_of_histogram(edges, contents):
    """Piecewise-linear CDF through the bin edges of a histogram: y starts at 0 on the first edge
    and ends at 1 on the last, so len(y) == len(edges).  (The reference's _pdf_to_cdf,
    chroma/detector.py:104-107, loses the leading 0 -- `[0.0] + array` adds instead of prepending --
    and its GPU side then reads one float past the end)"""
    x = np.array(edges, dtype=np.float64)
    y = np.zeros(len(x), dtype=np.float64)
    np.cumsum(contents, out=y[1:])
    return x, y / y[-1]


def gaussian_histogram(mean, rms, lo, hi, nbins):
    """Bin edges on [lo, hi] and the Gaussian density sampled at each bin's UPPER edge (the
    sampling chroma/detector.py:114-129 uses)."""
    edges = np.linspace(lo, hi, nbins + 1, endpoint=True)
    return edges, np.exp(-0.5 * ((edges[1:] - mean) / rms) ** 2)


class Detector(Geometry):
    def __init__(self, detector_material=None):
        Geometry.__init__(self, detector_material=detector_material)
        self.solid_id_to_channel_index = []
        self.channel_index_to_solid_id = []
        self.channel_index_to_channel_type = []
        self.channel_index_to_position = []
        # until told otherwise: no time smearing, unit charge (two-point CDFs a few ulps wide)
        self.time_cdf = (np.array([-1e-8, 1e-8]), np.array([0.0, 1.0]))
        self.charge_cdf = (np.array([0.999999999, 1.0]), np.array([0.0, 1.0]))

    # ---- solids and channels
    def add_solid(self, solid, rotation=None, displacement=None):
        """A passive solid; returns its solid id."""
        solid_id = Geometry.add_solid(self, solid, rotation=rotation, displacement=displacement)
        self.solid_id_to_channel_index.append(PASSIVE)
        return solid_id

    def _open_channel(self, solid_id, channel_type, position):
        channel = len(self.channel_index_to_solid_id)
        self.solid_id_to_channel_index[solid_id] = channel
        self.channel_index_to_solid_id.append(solid_id)
        self.channel_index_to_channel_type.append(channel if channel_type is None else channel_type)
        self.channel_index_to_position.append(position)
        return channel

    def add_pmt(self, pmt, rotation=None, displacement=None, channel_type=None):
        """A solid that is read out: the next free channel index is bound to it.  Returns the ids
        as a dict (solid_id, channel_index, channel_type)."""
        solid_id = self.add_solid(pmt, rotation=rotation, displacement=displacement)
        channel = self._open_channel(solid_id, channel_type, displacement)
        return {'solid_id': solid_id, 'channel_index': channel,
                'channel_type': self.channel_index_to_channel_type[channel]}

    def num_channels(self):
        return len(self.channel_index_to_channel_type)

    # ---- response
    _pdf_to_cdf = staticmethod(cdf_of_histogram)

    def set_time_dist(self, bin_edges, bin_contents):
        self.time_cdf = cdf_of_histogram(bin_edges, bin_contents)

    def set_charge_dist(self, bin_edges, bin_contents):
        self.charge_cdf = cdf_of_histogram(bin_edges, bin_contents)

    def set_time_dist_gaussian(self, rms, lo, hi, nsamples=50):
        self.set_time_dist(*gaussian_histogram(0.0, rms, lo, hi, nsamples))

    def set_charge_dist_gaussian(self, mean, rms, lo, hi, nsamples=50):
        self.set_charge_dist(*gaussian_histogram(mean, rms, lo, hi, nsamples))

    def flatten(self, dedupe_vertices=True):
        """Geometry.flatten plus the channel maps as int32 arrays (what the GPU side uploads)."""
        for name in ('solid_id_to_channel_index', 'channel_index_to_solid_id', 'channel_index_to_channel_type'):
            setattr(self, name, np.asarray(getattr(self, name), dtype=np.int32))
        Geometry.flatten(self, dedupe_vertices=dedupe_vertices)
