"""GPUDetector: GPUGeometry plus channel map and response CDFs
(role of chroma/gpu/detector.py:14-39)."""
import numpy as np

from .. import _lib
from ..gpuarray import DeviceArray
from .geometry import GPUGeometry


def cdf_arrays(cdf, what='cdf'):
    """(x, y) float32 arrays of a response CDF, equally long.

    The reference's own Detector._pdf_to_cdf (chroma/detector.py:104-107) builds y as
    ``np.array([0.0] + bin_contents.cumsum())``, which ADDS 0.0 instead of prepending it: y comes
    out one entry shorter than x, and the reference then uploads len(x) as the length of both
    (gpu/detector.py:22-38), so its sampler reads one float past the end of y.  Such a pair is
    completed here with the missing leading 0 (what the docstring there describes); any other
    length mismatch is an error."""
    x = np.ascontiguousarray(cdf[0], dtype=np.float32)
    y = np.ascontiguousarray(cdf[1], dtype=np.float32)
    if len(y) == len(x) - 1:
        y = np.concatenate([np.zeros(1, dtype=np.float32), y])
    if len(x) != len(y) or len(x) < 2:
        raise ValueError('%s: x and y need the same length >= 2 (got %d and %d)' % (what, len(x), len(y)))
    return x, y


class GPUDetector(GPUGeometry):
    def __init__(self, detector, wavelengths=None, print_usage=False):
        GPUGeometry.__init__(self, detector, wavelengths=wavelengths, print_usage=False)
        s2c = np.ascontiguousarray(detector.solid_id_to_channel_index, dtype=np.int32)
        tx, ty = cdf_arrays(detector.time_cdf, 'time_cdf')
        qx, qy = cdf_arrays(detector.charge_cdf, 'charge_cdf')
        self.nchannels = detector.num_channels()
        # charge quantum: cdf_x[-1] / 2^16 (chroma/gpu/detector.py:39)
        self.charge_unit = np.float32(detector.charge_cdf[0][-1] / 2 ** 16)
        _lib.check(_lib.lib().cb_detector_attach(
            self.handle, s2c.ctypes.data, len(s2c), int(self.nchannels), tx.ctypes.data, ty.ctypes.data,
            len(tx), qx.ctypes.data, qy.ctypes.data, len(qx), float(self.charge_unit)))
        self._refresh_views()
        info = self._info
        view = lambda p, n, dt: DeviceArray(n, dt, _alloc=self, _ptr=p)
        self.solid_id_to_channel_index_gpu = view(info.solid_id_to_channel_index, len(s2c), np.int32)
        self.time_cdf_x_gpu = view(info.time_cdf_x, len(tx), np.float32)
        self.time_cdf_y_gpu = view(info.time_cdf_y, len(ty), np.float32)
        self.charge_cdf_x_gpu = view(info.charge_cdf_x, len(qx), np.float32)
        self.charge_cdf_y_gpu = view(info.charge_cdf_y, len(qy), np.float32)
        self.detector_gpu = self.handle
        if print_usage:
            self.print_device_usage()
