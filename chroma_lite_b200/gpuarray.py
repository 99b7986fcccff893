"""Minimal GPUArray-like device array over the C ABI's raw allocations.

Covers what the reference's hot-path host code uses from pycuda.gpuarray
(SURVEY App. C): empty/zeros/to_gpu/zeros_like/ones_like, .get/.set/.fill,
.gpudata/.size/.dtype/.nbytes/len and contiguous 1-D slicing.
"""
import ctypes as C
import threading
import numpy as np

from . import _lib


class vec(object):
    """Stand-ins for pycuda.gpuarray.vec dtypes used by the reference."""
    float3 = np.dtype([('x', np.float32), ('y', np.float32), ('z', np.float32)])
    uint3 = np.dtype([('x', np.uint32), ('y', np.uint32), ('z', np.uint32)])
    uint4 = np.dtype([('x', np.uint32), ('y', np.uint32), ('z', np.uint32), ('w', np.uint32)])
    float4 = np.dtype([('x', np.float32), ('y', np.float32), ('z', np.float32), ('w', np.float32)])

    @staticmethod
    def make_float3(x, y, z):
        a = np.zeros((), dtype=vec.float3)
        a['x'], a['y'], a['z'] = x, y, z
        return a


class _Pool(object):
    """Size-bucketed cache of freed device blocks (cudaMalloc/cudaFree cost
    milliseconds and synchronise; a simulation allocates the same bank sizes for
    every batch).  Blocks are rounded up to a power of two below 16 MiB (hit lists
    change size from event to event) and to 2 MiB multiples above."""
    LIMIT = 16 << 30

    def __init__(self):
        self.free, self.cached, self.lock = {}, 0, threading.Lock()

    @staticmethod
    def bucket(nbytes):
        n = max(int(nbytes), 256)
        if n <= (16 << 20):
            return 1 << (n - 1).bit_length()
        g = 2 << 20
        return (n + g - 1) // g * g

    def take(self, size):
        with self.lock:
            lst = self.free.get(size)
            if lst:
                self.cached -= size
                return lst.pop()
        return None

    def give(self, ptr, size):
        with self.lock:
            if self.cached + size <= self.LIMIT:
                self.free.setdefault(size, []).append(ptr)
                self.cached += size
                return True
        return False

    def release_all(self):
        with self.lock:
            blocks = [p for lst in self.free.values() for p in lst]
            self.free, self.cached = {}, 0
        for p in blocks:
            _lib._lib.cb_free(C.c_void_p(p))

    def reserve(self, nbytes, count):
        """Make sure `count` free blocks of this size are cached: a pipeline allocates them NOW, while
        nothing is running, instead of meeting a cudaMalloc (which waits for the kernels in flight and can
        take milliseconds) the first time it happens to hold one block more than ever before."""
        size = self.bucket(nbytes)
        made = 0
        while True:
            with self.lock:
                if len(self.free.get(size, ())) >= count or self.cached + size > self.LIMIT:
                    return made
            p = C.c_void_p()
            if _lib.lib().cb_malloc(size, C.byref(p)) != 0:
                return made
            with self.lock:
                self.free.setdefault(size, []).append(p.value)
                self.cached += size
            made += 1


_pool = _Pool()


class _HostPool(object):
    """Page-locked host buffers for read-backs.  A device-to-host copy into a freshly allocated
    pageable array runs at ~4 GB/s on a B200 host (first-touch page faults under the driver's
    staging copy) against ~55 GB/s into page-locked memory, and cudaMallocHost itself costs a
    millisecond, so results are returned in pooled page-locked buffers: the buffer goes back to
    the pool when the last NumPy view of it is garbage-collected."""
    LIMIT = 2 << 30            # cached (unused) bytes kept
    LARGEST = 1 << 30          # bigger read-backs use a pageable array

    def __init__(self):
        self.free, self.cached, self.lock = {}, 0, threading.Lock()

    def array(self, shape, dtype):
        """Uninitialised NumPy array of the given shape in page-locked memory, or None."""
        dtype = np.dtype(dtype)
        count = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
        nbytes = count * dtype.itemsize
        if nbytes == 0 or nbytes > self.LARGEST or _lib._lib is None:
            return None
        size = _Pool.bucket(nbytes)
        ptr = None
        with self.lock:
            lst = self.free.get(size)
            if lst:
                ptr = lst.pop()
                self.cached -= size
        if ptr is None:
            p = C.c_void_p()
            if _lib._lib.cb_host_alloc(size, C.byref(p)) != 0:
                return None
            ptr = p.value
        buf = (C.c_char * size).from_address(ptr)
        import weakref
        weakref.finalize(buf, self._give, ptr, size)
        return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)

    def _give(self, ptr, size):
        try:
            with self.lock:
                if self.cached + size <= self.LIMIT:
                    self.free.setdefault(size, []).append(ptr)
                    self.cached += size
                    return
            if _lib._lib is not None:
                _lib._lib.cb_host_free(C.c_void_p(ptr))
        except Exception:
            pass

    def release_all(self):
        with self.lock:
            blocks = [p for lst in self.free.values() for p in lst]
            self.free, self.cached = {}, 0
        for p in blocks:
            _lib._lib.cb_host_free(C.c_void_p(p))


_host_pool = _HostPool()


def host_result(shape, dtype):
    """Array for a device-to-host read-back: pooled page-locked memory when possible."""
    if np.isscalar(shape):
        shape = (int(shape),)
    out = _host_pool.array(tuple(int(x) for x in shape), dtype)
    return out if out is not None else np.empty(shape, dtype=dtype)


def reserve(shape, dtype, count):
    """Pre-populate the device block cache with `count` blocks that fit an array of this shape."""
    n = int(np.prod(shape, dtype=np.int64)) if not np.isscalar(shape) else int(shape)
    return _pool.reserve(max(n * np.dtype(dtype).itemsize, 16), int(count))


def empty_cache():
    """Return every cached device block and page-locked result buffer to the driver."""
    if _lib._lib is not None:
        _pool.release_all()
        _host_pool.release_all()


class _Allocation(object):
    """Owns one device allocation; goes back to the pool when the last view dies."""
    def __init__(self, nbytes):
        self.size = _Pool.bucket(nbytes)
        ptr = _pool.take(self.size)
        if ptr is None:
            p = C.c_void_p()
            rc = _lib.lib().cb_malloc(self.size, C.byref(p))
            if rc != 0:                      # out of memory: drop the cache and retry once
                _pool.release_all()
                _lib.check(_lib.lib().cb_malloc(self.size, C.byref(p)))
            ptr = p.value
        self.ptr = ptr
        self.nbytes = int(nbytes)

    def __del__(self):
        try:
            if self.ptr and _lib._lib is not None:
                if not _pool.give(self.ptr, self.size):
                    _lib._lib.cb_free(C.c_void_p(self.ptr))
        except Exception:
            pass
        self.ptr = None


class DeviceArray(object):
    def __init__(self, shape, dtype, _alloc=None, _ptr=None):
        if np.isscalar(shape):
            shape = (int(shape),)
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.size = int(np.prod(self.shape)) if len(self.shape) else 1
        self.nbytes = self.size * self.dtype.itemsize
        if _alloc is None and _ptr is None:
            self._alloc = _Allocation(max(self.nbytes, 16))
            self.ptr = self._alloc.ptr
        else:
            self._alloc = _alloc      # keeps the owner alive (None for foreign memory)
            self.ptr = _ptr

    # -- pycuda-compatible surface
    @property
    def gpudata(self):
        return self.ptr

    def __int__(self):
        return int(self.ptr)

    def __len__(self):
        return self.shape[0] if self.shape else 1

    @property
    def __cuda_array_interface__(self):
        # lets torch.as_tensor(..., device='cuda') alias the memory (NCCL plumbing)
        if self.dtype.fields is not None:
            raise TypeError('structured dtype cannot be exported; use .view()')
        return {'shape': self.shape, 'typestr': self.dtype.str, 'data': (int(self.ptr), False),
                'version': 2}

    def view(self, dtype):
        dtype = np.dtype(dtype)
        assert self.nbytes % dtype.itemsize == 0
        return DeviceArray(self.nbytes // dtype.itemsize, dtype, _alloc=self._alloc, _ptr=self.ptr)

    def get(self):
        out = host_result(self.shape, self.dtype)
        if self.nbytes:
            _lib.check(_lib.lib().cb_memcpy_d2h(out.ctypes.data, self.ptr, self.nbytes))
        return out

    def set(self, arr):
        arr = np.ascontiguousarray(arr)
        if arr.nbytes != self.nbytes:
            raise ValueError('size mismatch in DeviceArray.set: %d != %d bytes' % (arr.nbytes, self.nbytes))
        if self.nbytes:
            _lib.check(_lib.lib().cb_memcpy_h2d(self.ptr, arr.ctypes.data, self.nbytes))
        return self

    def copy_from_device(self, other, nbytes=None):
        n = self.nbytes if nbytes is None else int(nbytes)
        _lib.check(_lib.lib().cb_memcpy_d2d(self.ptr, int(other), n))
        return self

    def fill(self, value):
        if self.dtype.itemsize != 4 or self.dtype.fields is not None:
            raise TypeError('fill() supports 4-byte scalar dtypes')
        bits = int(np.array(value, dtype=self.dtype).view(np.uint32))
        _lib.check(_lib.lib().cb_memset32(self.ptr, bits, self.size))
        return self

    def __getitem__(self, key):
        if not isinstance(key, slice):
            raise TypeError('DeviceArray supports contiguous slices only')
        start, stop, step = key.indices(len(self))
        if step != 1:
            raise ValueError('DeviceArray slices must be contiguous')
        n = max(0, stop - start)
        return DeviceArray(n, self.dtype, _alloc=self._alloc or self, _ptr=self.ptr + start * self.dtype.itemsize)


GPUArray = DeviceArray


def empty(shape, dtype):
    return DeviceArray(shape, dtype)


def zeros(shape, dtype):
    a = DeviceArray(shape, dtype)
    if a.nbytes:
        _lib.check(_lib.lib().cb_memset32(a.ptr, 0, (a.nbytes + 3) // 4))
    return a


def to_gpu(arr):
    arr = np.ascontiguousarray(arr)
    return DeviceArray(arr.shape, arr.dtype).set(arr)


def zeros_like(other, dtype=None):
    return zeros(other.shape, dtype or other.dtype)


def ones_like(other, dtype=None):
    a = DeviceArray(other.shape, dtype or other.dtype)
    a.fill(1)
    return a


def empty_like(other):
    return DeviceArray(other.shape, other.dtype)
