"""Boundary utilities (role of chroma/gpu/tools.py): context creation, RNG
pool, launch chunking helpers, float3/uint3 views."""
import ctypes as C
import numpy as np

from .. import _lib
from ..gpuarray import vec


def create_cuda_context(device_id=None):
    """Bind this process to one GPU (chroma/gpu/tools.py:182-203).  Returns a
    small object with pop()/synchronize() so Simulation's lifecycle code works."""
    dev = _lib.init(device_id)
    return _Context(dev)


class _Context(object):
    def __init__(self, device):
        self.device = device

    def synchronize(self):
        _lib.check(_lib.lib().cb_synchronize())

    def pop(self):
        pass

    def push(self):
        pass


class RNGStates(object):
    """Pool of XORWOW states; state i == curand_init(seed, first_stream + i, 0)
    (get_rng_states, chroma/gpu/tools.py:136-145, where first_stream is always 0).

    ``first_stream`` is what makes multi-GPU runs independent of the partition: give every rank
    the pool of ITS photons (first_stream = global index of its first photon) and hand each
    event the window of its own photons with ``view()``; stream id == global photon index."""

    def __init__(self, size, seed=1, offset=0, first_stream=0, _view_of=None, _handle=None):
        self.size = int(size)
        self.seed = seed
        self.first_stream = int(first_stream)
        self._parent = _view_of              # keeps the owning pool alive
        if _handle is not None:
            self.handle = _handle
            return
        h = C.c_uint64()
        _lib.check(_lib.lib().cb_rng_create_streams(self.size, int(seed) & 0xFFFFFFFFFFFFFFFF, self.first_stream,
                                                    int(offset), C.byref(h)))
        self.handle = h.value

    def view(self, first, count):
        """Non-owning window [first, first + count) of this pool."""
        h = C.c_uint64()
        _lib.check(_lib.lib().cb_rng_view(self.handle, int(first), int(count), C.byref(h)))
        return RNGStates(count, seed=self.seed, first_stream=self.first_stream + int(first), _view_of=self, _handle=h.value)

    def __len__(self):
        return self.size

    def get(self, first=0, count=None):
        """State words {d, v0..v4} as uint32 (count, 6)."""
        count = self.size - first if count is None else count
        out = np.empty((count, 6), dtype=np.uint32)
        _lib.check(_lib.lib().cb_rng_download(self.handle, int(first), int(count), out.ctypes.data))
        return out

    def __del__(self):
        try:
            if self.handle and _lib._lib is not None:
                _lib._lib.cb_rng_destroy(self.handle)
        except Exception:
            pass
        self.handle = 0


def get_rng_states(size, seed=1, first_stream=0):
    "Return `size` number of CUDA random number generator states."
    return RNGStates(size, seed=seed, first_stream=first_stream)


def to_float3(arr):
    "(N,3) array -> float3 structured array (chroma/gpu/tools.py:147-151)."
    arr = np.ascontiguousarray(arr, dtype=np.float32)
    return arr.view(vec.float3)[:, 0]


def to_uint3(arr):
    arr = np.ascontiguousarray(arr, dtype=np.uint32)
    return arr.view(vec.uint3)[:, 0]


def chunk_iterator(nelements, nthreads_per_block=64, max_blocks=1024):
    """(first_index, elements_this_iteration, nblocks_this_iteration) tuples
    (chroma/gpu/tools.py:159-180).

    >>> list(chunk_iterator(300, 32, 2))
    [(0, 64, 2), (64, 64, 2), (128, 64, 2), (192, 64, 2), (256, 44, 2)]
    """
    first = 0
    while first < nelements:
        left = nelements - first
        blocks = min(max_blocks, -(-left // nthreads_per_block))
        n = min(left, blocks * nthreads_per_block)
        yield (first, n, blocks)
        first += n


def format_size(size):
    for lim, div, suf in ((1e3, 1, ' '), (1e6, 1e3, 'K'), (1e9, 1e6, 'M')):
        if size < lim:
            return '%.1f%s' % (size / div, suf)
    return '%.1f%s' % (size / 1e9, 'G')


def format_array(name, array):
    return '%-15s %6s %6s' % (name, format_size(len(array)), format_size(array.nbytes))


# ------------------------------------------------------------------ page-locked host arrays
# role of pycuda.driver.pagelocked_empty / chroma.gpu.tools.mapped_empty
# (chroma/gpu/tools.py:244-283): numpy arrays in page-locked host memory, which the
# copy engines read at PCIe speed instead of being staged through the driver's bounce
# buffers.  Event arrays built in them upload ~3x faster than pageable numpy arrays.
import weakref


def pagelocked_empty(shape, dtype, write_combined=False, **_ignored):
    """Page-locked host array (pycuda.driver.pagelocked_empty).  write_combined=True (pycuda's
    host_alloc_flags.WRITECOMBINED): for buffers the CPU only writes and the GPU only reads."""
    dtype = np.dtype(dtype)
    shape = (int(shape),) if np.isscalar(shape) else tuple(int(x) for x in shape)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    ptr = C.c_void_p()
    _lib.check(_lib.lib().cb_host_alloc_flags(max(nbytes, 16), int(bool(write_combined)), C.byref(ptr)))
    buf = (C.c_char * max(nbytes, 16)).from_address(ptr.value)
    weakref.finalize(buf, _lib.lib().cb_host_free, C.c_void_p(ptr.value))
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)


def pagelocked_zeros(shape, dtype, **_ignored):
    a = pagelocked_empty(shape, dtype)
    a[...] = 0
    return a


def pagelocked_copy(arr, write_combined=False):
    arr = np.asarray(arr)
    a = pagelocked_empty(arr.shape, arr.dtype, write_combined=write_combined)
    a[...] = arr
    return a


mapped_empty = pagelocked_empty
mapped_zeros = pagelocked_zeros


def mapped_empty_like(other, **_ignored):
    return pagelocked_empty(other.shape, other.dtype)


def mapped_zeros_like(other, **_ignored):
    return pagelocked_zeros(other.shape, other.dtype)


def pin_photons(photons, write_combined=False):
    """A copy of an event.Photons whose arrays live in page-locked host memory."""
    from .. import event
    f = ('pos', 'dir', 'pol', 'wavelengths', 't', 'last_hit_triangles', 'flags', 'weights', 'evidx')
    out = event.Photons.__new__(event.Photons)
    for k in f:
        setattr(out, k, pagelocked_copy(getattr(photons, k), write_combined=write_combined))
    out.channel = np.asarray(getattr(photons, 'channel', np.zeros(len(photons), dtype=np.uint32)))
    return out
