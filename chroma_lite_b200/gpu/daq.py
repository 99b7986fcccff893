"""GPUDaq / GPUChannels (role of chroma/gpu/daq.py:8-101)."""
import ctypes as C
import numpy as np

from .. import _lib, event
from ..gpuarray import DeviceArray


class GPUChannels(object):
    def __init__(self, t, q, flags, ndaq=1, stride=None):
        self.t, self.q, self.flags, self.ndaq = t, q, flags, ndaq
        self.stride = len(t) if stride is None else stride

    def iterate_copies(self):
        for i in range(self.ndaq):
            w = slice(i * self.stride, (i + 1) * self.stride)
            yield GPUChannels(self.t[w], self.q[w], self.flags[w])

    def get(self):
        t = self.t.get()
        q = self.q.get()
        # channels with a small enough time were hit (gpu/daq.py:26-32)
        return event.Channels(t < 1e8, t, q, self.flags.get())

    def __len__(self):
        return self.t.size


class GPUDaq(object):
    def __init__(self, gpu_detector, ndaq=1):
        assert gpu_detector.nchannels > 0, "Geometry has no detectors, DAQ can't be initialized."
        h = C.c_uint64()
        _lib.check(_lib.lib().cb_daq_create(gpu_detector.handle, int(ndaq), C.byref(h)))
        self.handle = h.value
        self.ndaq = ndaq
        self.stride = gpu_detector.nchannels
        self.gpu_detector = gpu_detector
        p = [C.c_void_p() for _ in range(5)]
        n = C.c_uint64()
        _lib.check(_lib.lib().cb_daq_pointers(self.handle, *[C.byref(x) for x in p], C.byref(n)))
        view = lambda ptr, dt: DeviceArray(n.value, dt, _alloc=self, _ptr=ptr.value)
        self.earliest_time_gpu = view(p[0], np.float32)
        self.channel_q_gpu = view(p[1], np.float32)
        self.channel_history_gpu = view(p[2], np.uint32)
        self.earliest_time_int_gpu = view(p[3], np.uint32)
        self.channel_q_int_gpu = view(p[4], np.uint32)

    def begin_acquire(self, nthreads_per_block=64):
        _lib.check(_lib.lib().cb_daq_begin_acquire(self.handle))

    def acquire(self, gpuphotons, rng_states, nthreads_per_block=64, max_blocks=1024, start_photon=None,
                nphotons=None, weight=1.0):
        start_photon = 0 if start_photon is None else start_photon
        nphotons = len(gpuphotons.pos) - start_photon if nphotons is None else nphotons
        bank = gpuphotons._bank()
        _lib.check(_lib.lib().cb_daq_acquire(self.handle, C.byref(bank), rng_states.handle, int(nthreads_per_block),
                                             int(max_blocks), int(start_photon), int(nphotons), float(weight)))

    def acquire_async(self, gpuphotons, rng_states, nthreads_per_block=64, max_blocks=1024, start_photon=None,
                      nphotons=None, weight=1.0, begin=True, finalize=True):
        """begin_acquire + acquire + end_acquire of one acquisition, only enqueued on the library
        stream (the simulation pipeline's form); returns the GPUChannels, valid once a marker recorded
        behind this call has been waited for."""
        start_photon = 0 if start_photon is None else start_photon
        nphotons = len(gpuphotons.pos) - start_photon if nphotons is None else nphotons
        bank = gpuphotons._bank()
        _lib.check(_lib.lib().cb_daq_acquire_async(self.handle, C.byref(bank), rng_states.handle, int(nthreads_per_block),
                                                   int(max_blocks), int(start_photon), int(nphotons), float(weight),
                                                   int(bool(begin)), int(bool(finalize))))
        return GPUChannels(self.earliest_time_gpu, self.channel_q_gpu, self.channel_history_gpu, self.ndaq, self.stride)

    def end_acquire(self, nthreads_per_block=64):
        _lib.check(_lib.lib().cb_daq_end_acquire(self.handle))
        return GPUChannels(self.earliest_time_gpu, self.channel_q_gpu, self.channel_history_gpu, self.ndaq, self.stride)

    def fold(self, other, wait=True):
        """Merge another GPUDaq's accumulators into this one on the device (MIN time, SUM charge, OR
        history): per-event acquisitions into run-level accumulators.  wait=False only enqueues the
        merge (ordered on the library stream, so a later allreduce() / end_acquire() sees it)."""
        if wait:
            _lib.check(_lib.lib().cb_daq_fold(self.handle, other.handle))
        else:
            _lib.check(_lib.lib().cb_daq_fold_async(self.handle, other.handle))

    def allreduce(self):
        """Combine the accumulators of every rank (MIN time, SUM charge, OR history) with one
        grouped NCCL all-reduce inside the library and convert to the float outputs; needs
        parallel.init_comm() first, and is end_acquire() on a single rank.  Collective."""
        _lib.check(_lib.lib().cb_daq_allreduce(self.handle))
        return GPUChannels(self.earliest_time_gpu, self.channel_q_gpu, self.channel_history_gpu, self.ndaq, self.stride)

    def finalize_reduced(self):
        """int accumulators -> float outputs (after a cross-GPU reduction)."""
        _lib.check(_lib.lib().cb_daq_finalize(self.handle))
        return GPUChannels(self.earliest_time_gpu, self.channel_q_gpu, self.channel_history_gpu, self.ndaq, self.stride)

    def __del__(self):
        try:
            if getattr(self, 'handle', 0) and _lib._lib is not None:
                _lib._lib.cb_daq_destroy(self.handle)
        except Exception:
            pass
        self.handle = 0
