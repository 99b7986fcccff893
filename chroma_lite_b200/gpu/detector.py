"""GPUDetector: GPUGeometry plus channel map and response CDFs
(role of chroma/gpu/detector.py:14-39)."""
import numpy as np

from .. import _lib
from ..gpuarray import DeviceArray
from .geometry import GPUGeometry


class GPUDetector(GPUGeometry):
    def __init__(self, detector, wavelengths=None, print_usage=False):
        GPUGeometry.__init__(self, detector, wavelengths=wavelengths, print_usage=False)
        s2c = np.ascontiguousarray(detector.solid_id_to_channel_index, dtype=np.int32)
        tx = np.ascontiguousarray(detector.time_cdf[0], dtype=np.float32)
        ty = np.ascontiguousarray(detector.time_cdf[1], dtype=np.float32)
        qx = np.ascontiguousarray(detector.charge_cdf[0], dtype=np.float32)
        qy = np.ascontiguousarray(detector.charge_cdf[1], dtype=np.float32)
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
