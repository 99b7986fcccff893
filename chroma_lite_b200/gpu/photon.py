"""GPUPhotons / GPUPhotonsSlice: device photon bank and its operations
(role of chroma/gpu/photon.py:13-415), implemented on the C ABI."""
import ctypes as C
import os
import sys
import numpy as np

from .. import _lib, event
from .. import gpuarray as ga
from .tools import to_float3

_FIELDS = ('pos', 'dir', 'pol', 'wavelengths', 't', 'last_hit_triangles', 'flags', 'weights', 'evidx')


def _resolve_nphotons(ph):
    try:
        return len(ph)
    except TypeError:
        pass
    true_n = getattr(ph, 'true_nphotons', None)
    if true_n is not None:
        return int(true_n)
    pos = getattr(ph, 'pos', None)
    if pos is not None:
        return len(pos)
    raise TypeError('Cannot determine photon count from object of type %r' % type(ph))


_pool = None


def _upload_threads():
    """Host threads that stage a bank's arrays concurrently (each with its own copy stream and
    spinning in its stream synchronisation).  Three when this process has the host to itself; one
    when several ranks share it (LOCAL_WORLD_SIZE / WORLD_SIZE of torchrun >= 4): a single stream
    still uploads a 130 MB event in ~3 ms, well inside the 7 ms GPU stage it overlaps with, and the
    ranks' spinning threads no longer outnumber the cores.  CHROMA_B200_UPLOAD_THREADS overrides."""
    import os
    if os.environ.get('CHROMA_B200_UPLOAD_THREADS'):
        return max(1, int(os.environ['CHROMA_B200_UPLOAD_THREADS']))
    ranks = int(os.environ.get('LOCAL_WORLD_SIZE', os.environ.get('WORLD_SIZE', '1')) or 1)
    return 1 if ranks >= 4 else 3


def _upload_pool():
    global _pool
    if _pool is None:
        import concurrent.futures
        _pool = concurrent.futures.ThreadPoolExecutor(max_workers=_upload_threads(), thread_name_prefix='cb-upload')
    return _pool


def _alloc_bank(n):
    return dict(pos=ga.empty(n, ga.vec.float3), dir=ga.empty(n, ga.vec.float3), pol=ga.empty(n, ga.vec.float3),
                wavelengths=ga.empty(n, np.float32), t=ga.empty(n, np.float32),
                last_hit_triangles=ga.empty(n, np.int32), flags=ga.empty(n, np.uint32),
                weights=ga.empty(n, np.float32), evidx=ga.empty(n, np.uint32))


def reserve_banks(nphotons, count):
    """Cache the device blocks of `count` photon banks of `nphotons` photons and of their packed hit blocks
    (GPUPhotons.flat_hits_async) ahead of time: a pipeline that holds several batches in flight then never
    allocates while kernels run."""
    ga.reserve(3 * nphotons, np.float32, 3 * count)   # pos, dir, pol
    ga.reserve(nphotons, np.float32, 6 * count)       # wavelengths, t, last_hit_triangles, flags, weights, evidx
    ga.reserve(max(int(nphotons), 1) * 16 + 1, np.uint32, max(2, count - 1))


class GPUPhotons(object):
    def __init__(self, photons, ncopies=1, copy_flags=True, copy_triangles=True, copy_weights=True, evidx_value=None):
        """Load ``photons`` onto the GPU, replicating ``ncopies`` times
        (chroma/gpu/photon.py:14-116).  ``evidx_value`` (extension): every photon belongs to this event
        of its batch -- the array is filled on the device instead of being uploaded."""
        nphotons = _resolve_nphotons(photons)
        total = nphotons * ncopies
        self.pos = ga.empty(total, ga.vec.float3)
        self.dir = ga.empty(total, ga.vec.float3)
        self.pol = ga.empty(total, ga.vec.float3)
        self.wavelengths = ga.empty(total, np.float32)
        self.t = ga.empty(total, np.float32)
        self.last_hit_triangles = ga.empty(total, np.int32)
        self.flags = ga.empty(total, np.uint32)
        self.weights = ga.empty(total, np.float32)
        # the reference allocates evidx for nphotons only although photon_duplicate
        # writes all copies; allocate the full size
        self.evidx = ga.empty(total, np.uint32)

        wanted = {'pos': True, 'dir': True, 'pol': True, 'wavelengths': True, 't': True, 'evidx': evidx_value is None,
                  'last_hit_triangles': copy_triangles, 'flags': copy_flags, 'weights': copy_weights}
        on_host = nphotons > 0 and not any(isinstance(getattr(photons, f, None), ga.DeviceArray) for f, w in wanted.items() if w)
        if on_host:
            # host arrays: the whole bank in one library call (copies and default fills on this thread's copy
            # stream, one wait, the interpreter lock released throughout)
            host, keep = _lib.CbPhotonBank(), []
            self.h2d_bytes = 0
            skip_zero = nphotons >= 65536 and os.environ.get('CHROMA_B200_SKIP_ZERO', '1') != '0'
            for f in _FIELDS:
                if not wanted[f]:
                    continue
                if f in ('pos', 'dir', 'pol'):
                    a = to_float3(np.asarray(getattr(photons, f)))
                else:
                    a = np.ascontiguousarray(getattr(photons, f), dtype=getattr(self, f).dtype)
                if a.nbytes != nphotons * getattr(self, f).dtype.itemsize:
                    raise ValueError('photon field %s has %d bytes for %d photons' % (f, a.nbytes, nphotons))
                # flags and times of freshly generated photons are usually all zero: 0.3 ms of host time to
                # find out, against 4 bytes per photon over PCIe; the device array is zeroed in place instead
                if skip_zero and f in ('flags', 't') and a.view(np.uint32).max() == 0:
                    continue
                keep.append(a)
                self.h2d_bytes += a.nbytes
                setattr(host, f, a.ctypes.data)
            dst = self._bank(0, nphotons)
            _lib.check(_lib.lib().cb_photon_bank_upload(C.byref(dst), C.byref(host), int(nphotons),
                                                        int(evidx_value or 0)))
            del keep
        else:
            if not copy_triangles:
                self.last_hit_triangles.fill(-1)
            if not copy_flags:
                self.flags.fill(0)
            if not copy_weights:
                self.weights.fill(1.0)

        def put_vec(dest, source):
            if isinstance(source, ga.DeviceArray):
                dest[:nphotons].copy_from_device(source, min(len(source), nphotons) * 12)
            else:
                dest[:nphotons].set(to_float3(np.asarray(source)))

        def put(dest, source, dtype):
            if isinstance(source, ga.DeviceArray):
                dest[:nphotons].copy_from_device(source, min(len(source), nphotons) * 4)
            else:
                dest[:nphotons].set(np.asarray(source, dtype=dtype))

        if nphotons and not on_host:
            jobs = [(put_vec, self.pos, photons.pos), (put_vec, self.dir, photons.dir), (put_vec, self.pol, photons.pol),
                    (put, self.wavelengths, photons.wavelengths, np.float32), (put, self.t, photons.t, np.float32)]
            if evidx_value is None:
                jobs.append((put, self.evidx, photons.evidx, np.uint32))
            else:
                jobs.append((lambda dest, value: dest[:nphotons].fill(np.uint32(value)), self.evidx, evidx_value))
            if copy_triangles:
                jobs.append((put, self.last_hit_triangles, photons.last_hit_triangles, np.int32))
            if copy_flags:
                jobs.append((put, self.flags, photons.flags, np.uint32))
            if copy_weights:
                jobs.append((put, self.weights, photons.weights, np.float32))
            if nphotons >= 200000:
                # large banks: stage the arrays from a few host threads at once (each
                # thread has its own copy stream; the C calls release the GIL)
                list(_upload_pool().map(lambda j: j[0](*j[1:]), jobs))
            else:
                for j in jobs:
                    j[0](*j[1:])

        self.true_nphotons = getattr(photons, 'true_nphotons', nphotons)
        self.ncopies = ncopies
        if ncopies > 1 and nphotons:
            bank = self._bank()
            _lib.check(_lib.lib().cb_photon_duplicate(C.byref(bank), int(nphotons), int(ncopies)))

    # ------------------------------------------------------------------
    def _bank(self, start=0, count=None):
        n = len(self.pos) - start if count is None else count
        b = _lib.CbPhotonBank()
        for f in _FIELDS:
            arr = getattr(self, f)
            setattr(b, f, arr.ptr + start * arr.dtype.itemsize if arr is not None else None)
        b.n = int(n)
        return b

    def get(self):
        n = len(self.pos)
        pos = self.pos.get().view(np.float32).reshape((n, 3))
        dir = self.dir.get().view(np.float32).reshape((n, 3))
        pol = self.pol.get().view(np.float32).reshape((n, 3))
        return event.Photons(pos, dir, pol, self.wavelengths.get(), self.t.get(), self.last_hit_triangles.get(),
                             self.flags.get(), self.weights.get(), self.evidx.get())

    def get_hits(self, *args, **kwargs):
        """dict channel -> Photons detected by that channel."""
        flat_hits = self.get_flat_hits(*args, **kwargs)
        return {int(chan): flat_hits[flat_hits.channel == chan] for chan in np.unique(flat_hits.channel)}

    def get_flat_hits(self, gpu_detector, target_flag=(0x1 << 2), nthreads_per_block=256, max_blocks=1024,
                      start_photon=None, nphotons=None, no_map=False):
        """Photons with ``target_flag`` set that ended on a triangle of a solid mapped
        to a channel; ``.channel`` holds the channel index (gpu/photon.py:141-209)."""
        import time as _time
        _t0 = _time.perf_counter()
        lib = _lib.lib()
        start_photon = 0 if start_photon is None else start_photon
        nphotons = self.pos.size - start_photon if nphotons is None else nphotons
        src = self._bank()
        count = C.c_uint32()
        _lib.check(lib.cb_count_photon_hits(C.byref(src), int(start_photon), int(nphotons), int(target_flag),
                                            gpu_detector.handle, C.byref(count)))
        n = count.value
        # the ten result arrays live in ONE device block and come back in ONE copy into a pooled
        # page-locked buffer (the reference: ten gpuarray allocations and ten .get() calls)
        words = (3, 3, 3, 1, 1, 1, 1, 1, 1, 1)             # 4-byte words per photon: _FIELDS + channel
        block = ga.empty(max(n, 1) * sum(words), np.uint32)
        offs = np.concatenate([[0], np.cumsum(words)]) * max(n, 1)
        if n:
            dst = _lib.CbPhotonBank()
            for f, o in zip(_FIELDS, offs):
                setattr(dst, f, block.ptr + 4 * int(o))
            dst.n = n
            c2 = C.c_uint32()
            _lib.check(lib.cb_copy_photon_hits(C.byref(src), int(start_photon), int(nphotons), int(target_flag),
                                               gpu_detector.handle, C.byref(dst), block.ptr + 4 * int(offs[9]), C.byref(c2)))
            assert c2.value == n
        _t1 = _time.perf_counter()
        self.last_hit_timings = (_t1 - _t0,)
        host = block.get()
        part = lambda i, dt: host[int(offs[i]):int(offs[i]) + n * words[i]].view(dt)
        vec3 = lambda i: part(i, np.float32).reshape((n, 3))
        return event.Photons(vec3(0), vec3(1), vec3(2), part(3, np.float32), part(4, np.float32), part(5, np.int32),
                             part(6, np.uint32), part(7, np.float32), part(8, np.uint32), part(9, np.int32))

    def flat_hits_async(self, gpu_detector, target_flag=(0x1 << 2), start_photon=None, nphotons=None):
        """get_flat_hits in two halves for the simulation pipeline: this one only ENQUEUES the
        compaction behind whatever is running on the library stream and returns a PendingHits; its
        get() -- normally called by another thread while the GPU already works on the next batch --
        waits for the completion marker and reads the hits back in one copy."""
        lib = _lib.lib()
        start_photon = 0 if start_photon is None else start_photon
        nphotons = self.pos.size - start_photon if nphotons is None else nphotons
        src = self._bank()
        block = ga.empty(max(int(nphotons), 1) * 16 + 1, np.uint32)        # ten arrays of H hits + the count H
        count_ptr = block.ptr + 4 * max(int(nphotons), 1) * 16
        _lib.check(lib.cb_copy_photon_hits_async(C.byref(src), int(start_photon), int(nphotons), int(target_flag),
                                                 gpu_detector.handle, block.ptr, count_ptr))
        return PendingHits(block, count_ptr, self)

    def iterate_copies(self):
        for i in range(self.ncopies):
            w = slice(self.true_nphotons * i, self.true_nphotons * (i + 1))
            yield GPUPhotonsSlice(pos=self.pos[w], dir=self.dir[w], pol=self.pol[w], wavelengths=self.wavelengths[w],
                                  t=self.t[w], last_hit_triangles=self.last_hit_triangles[w], flags=self.flags[w],
                                  weights=self.weights[w], evidx=self.evidx[w])

    def propagate(self, gpu_geometry, rng_states, nthreads_per_block=256, max_blocks=1024, max_steps=10,
                  use_weights=False, scatter_first=0, track=False):
        """Propagate photons to termination or ``max_steps`` (gpu/photon.py:227-290).

        The whole step loop runs on the device in one call.  ``rng_states`` needs
        ``nthreads_per_block*max_blocks`` states; with a pool >= len(self) photon i
        uses stream i (replay contract); larger banks reuse the pool chunk by chunk.
        With ``track=True`` returns (step_photon_ids, step_photons) like the reference.
        """
        lib = _lib.lib()
        bank = self._bank()
        self.last_stats = _lib.CbPropagateStats()
        if not track:
            _lib.check(lib.cb_propagate(C.byref(bank), gpu_geometry.handle, rng_states.handle,
                                        int(nthreads_per_block), int(max_blocks), int(max_steps),
                                        int(bool(use_weights)), int(scatter_first), C.byref(self.last_stats)))
            return None
        # tracking: one step per call, snapshotting the still-alive photons
        n = self.pos.size
        ids = np.arange(n, dtype=np.uint32)
        step_photon_ids, step_photons = [ids], [self.copy_queue(ga.to_gpu(ids), n).get()]
        for step in range(max_steps):
            _lib.check(lib.cb_propagate(C.byref(bank), gpu_geometry.handle, rng_states.handle,
                                        int(nthreads_per_block), int(max_blocks), 1, int(bool(use_weights)),
                                        int(scatter_first), None))
            scatter_first = 0
            flags = self.flags.get()
            alive_prev = ids
            step_photon_ids.append(alive_prev)
            step_photons.append(self.copy_queue(ga.to_gpu(alive_prev), len(alive_prev)).get())
            ids = alive_prev[(flags[alive_prev] & event.TERMINAL_MASK) == 0]
            if len(ids) == 0:
                break
        return step_photon_ids, step_photons

    def copy_queue(self, queue_gpu, nphotons, nthreads_per_block=256, max_blocks=1024, start_photon=0):
        out = _alloc_bank(nphotons)
        if nphotons:
            src = self._bank()
            dst = _lib.CbPhotonBank()
            for f in _FIELDS:
                setattr(dst, f, out[f].ptr)
            dst.n = nphotons
            _lib.check(_lib.lib().cb_copy_photon_queue(C.byref(src), queue_gpu.ptr + 4 * start_photon,
                                                       int(nphotons), C.byref(dst)))
        return GPUPhotonsSlice(**out)

    def select(self, target_flag, nthreads_per_block=256, max_blocks=1024, start_photon=None, nphotons=None):
        """New bank with only the photons that have ``target_flag`` set
        (gpu/photon.py:321-371); photon order is preserved."""
        lib = _lib.lib()
        start_photon = 0 if start_photon is None else start_photon
        nphotons = self.pos.size - start_photon if nphotons is None else nphotons
        src = self._bank()
        count = C.c_uint32()
        _lib.check(lib.cb_count_photons(C.byref(src), int(start_photon), int(nphotons), int(target_flag), C.byref(count)))
        n = count.value
        out = _alloc_bank(n)
        if n:
            dst = _lib.CbPhotonBank()
            for f in _FIELDS:
                setattr(dst, f, out[f].ptr)
            dst.n = n
            c2 = C.c_uint32()
            _lib.check(lib.cb_copy_photons(C.byref(src), int(start_photon), int(nphotons), int(target_flag),
                                           C.byref(dst), C.byref(c2)))
            assert c2.value == n
        return GPUPhotonsSlice(**out)

    def __len__(self):
        return self.pos.size


class Marker(object):
    """Completion marker on the library stream (cb_event_*); reusable."""
    def __init__(self):
        h = C.c_uint64()
        _lib.check(_lib.lib().cb_event_create(C.byref(h)))
        self.handle = h.value

    def record(self):
        _lib.check(_lib.lib().cb_event_record(self.handle))
        return self

    def wait(self):
        _lib.check(_lib.lib().cb_event_wait(self.handle))

    def __del__(self):
        try:
            if getattr(self, 'handle', 0) and _lib._lib is not None:
                _lib._lib.cb_event_destroy(self.handle)
        except Exception:
            pass
        self.handle = 0


class PendingHits(object):
    """Hits being compacted on the device (GPUPhotons.flat_hits_async)."""
    WORDS = (3, 3, 3, 1, 1, 1, 1, 1, 1, 1)             # 4-byte words per hit: _FIELDS + channel

    def __init__(self, block, count_ptr, source):
        self.block, self.count_ptr, self.source = block, count_ptr, source      # `source` keeps the bank alive

    def get(self, marker=None, ready=False):
        """event.Photons of the hits (with .channel).  `marker` must have been recorded behind the
        compaction; ready=True: the caller has already waited for it (once for everything the batch
        enqueued); with neither, the whole device is synchronised."""
        if ready:
            pass
        elif marker is not None:
            marker.wait()
        else:
            _lib.check(_lib.lib().cb_synchronize())
        cnt = np.zeros(1, np.uint32)
        _lib.check(_lib.lib().cb_memcpy_d2h(cnt.ctypes.data, self.count_ptr, 4))
        n = int(cnt[0])
        host = self.block[:max(n, 1) * 16].get()
        self.block = self.source = None
        offs = np.concatenate([[0], np.cumsum(self.WORDS)]) * max(n, 1)
        part = lambda i, dt: host[int(offs[i]):int(offs[i]) + n * self.WORDS[i]].view(dt)
        vec3 = lambda i: part(i, np.float32).reshape((n, 3))
        return event.Photons(vec3(0), vec3(1), vec3(2), part(3, np.float32), part(4, np.float32), part(5, np.int32),
                             part(6, np.uint32), part(7, np.float32), part(8, np.uint32), part(9, np.int32))


class GPUPhotonsSlice(GPUPhotons):
    """A view (or freshly compacted bank) that behaves like GPUPhotons
    (gpu/photon.py:388-415)."""

    def __init__(self, pos, dir, pol, wavelengths, t, last_hit_triangles, flags, weights, evidx):
        self.pos, self.dir, self.pol = pos, dir, pol
        self.wavelengths, self.t = wavelengths, t
        self.last_hit_triangles, self.flags, self.weights, self.evidx = last_hit_triangles, flags, weights, evidx
        self.true_nphotons = len(pos)
        self.ncopies = 1
