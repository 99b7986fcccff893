"""Run the REFERENCE's own CUDA kernels (oracle/_ref/*.cubin, built from the
reference sources by oracle/Makefile) on the GPU.  TEST INFRASTRUCTURE ONLY.

PyCUDA is not installable here, so this is a direct CUDA-driver-API harness
(ctypes on libcuda) that replays the reference host code launch for launch:
  get_rng_states      chroma/gpu/tools.py:136-145  (init_rng, block 64)
  GPUGeometry structs chroma/gpu/geometry.py:44-520 (make_gpu_struct layouts of
                      geometry_types.h: Material 88 B, Surface 104 B, Geometry 96 B)
  GPUPhotons.propagate chroma/gpu/photon.py:240-290 (queues, chunk_iterator,
                      nsteps rule, 4-byte D2H per step)
  GPUDaq              chroma/gpu/daq.py:55-101
It shares the primary context with libchroma_b200 but no code or memory.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, '_ref')
_cu = None
_ctx_ready = False


class RefError(RuntimeError):
    pass


def _ck(rc, what=''):
    if rc != 0:
        raise RefError('CUDA driver error %d in %s' % (rc, what))


def available():
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in ('propagate.cubin', 'daq.cubin', 'ref_wrap.cubin'))


def cu():
    global _cu, _ctx_ready
    if _cu is None:
        _cu = C.CDLL('libcuda.so.1')
    if not _ctx_ready:
        _ck(_cu.cuInit(0), 'cuInit')
        dev = C.c_int()
        _ck(_cu.cuDeviceGet(C.byref(dev), int(os.environ.get('CHROMA_REF_DEVICE', '0'))), 'cuDeviceGet')
        ctx = C.c_void_p()
        _ck(_cu.cuDevicePrimaryCtxRetain(C.byref(ctx), dev), 'cuDevicePrimaryCtxRetain')
        _ck(_cu.cuCtxSetCurrent(ctx), 'cuCtxSetCurrent')
        _ctx_ready = True
    return _cu


class DevMem(object):
    def __init__(self, nbytes):
        self.nbytes = max(int(nbytes), 16)
        p = C.c_uint64()
        _ck(cu().cuMemAlloc_v2(C.byref(p), C.c_size_t(self.nbytes)), 'cuMemAlloc')
        self.ptr = p.value

    def __del__(self):
        try:
            if self.ptr and _cu is not None:
                _cu.cuMemFree_v2(C.c_uint64(self.ptr))
        except Exception:
            pass
        self.ptr = 0


def to_dev(arr):
    arr = np.ascontiguousarray(arr)
    m = DevMem(arr.nbytes)
    if arr.nbytes:
        _ck(cu().cuMemcpyHtoD_v2(C.c_uint64(m.ptr), arr.ctypes.data_as(C.c_void_p), C.c_size_t(arr.nbytes)), 'HtoD')
    m.dtype, m.shape = arr.dtype, arr.shape
    return m


def from_dev(m, dtype=None, shape=None, offset=0, nbytes=None):
    dtype = np.dtype(dtype or m.dtype)
    if shape is None:
        shape = m.shape
    out = np.empty(shape, dtype=dtype)
    n = out.nbytes if nbytes is None else nbytes
    if n:
        _ck(cu().cuMemcpyDtoH_v2(out.ctypes.data_as(C.c_void_p), C.c_uint64(m.ptr + offset), C.c_size_t(n)), 'DtoH')
    return out


def sync():
    _ck(cu().cuCtxSynchronize(), 'sync')


class Module(object):
    def __init__(self, name):
        path = os.path.join(REF_DIR, name)
        with open(path, 'rb') as f:
            self.image = f.read()
        self.mod = C.c_void_p()
        _ck(cu().cuModuleLoadData(C.byref(self.mod), self.image), 'cuModuleLoadData(%s)' % name)
        self.funcs = {}

    def func(self, name):
        if name not in self.funcs:
            f = C.c_void_p()
            _ck(cu().cuModuleGetFunction(C.byref(f), self.mod, name.encode()), 'cuModuleGetFunction(%s)' % name)
            self.funcs[name] = f
        return self.funcs[name]

    def launch(self, name, grid, block, *args):
        """args: ctypes scalars, DevMem (passed as pointer) or raw ints (pointers)."""
        vals = []
        for a in args:
            if isinstance(a, DevMem):
                vals.append(C.c_uint64(a.ptr))
            elif a is None:
                vals.append(C.c_uint64(0))
            elif isinstance(a, int):
                vals.append(C.c_uint64(a))
            else:
                vals.append(a)
        ptrs = (C.c_void_p * len(vals))(*[C.cast(C.byref(v), C.c_void_p) for v in vals])
        _ck(cu().cuLaunchKernel(self.func(name), int(grid), 1, 1, int(block), 1, 1, 0, None, ptrs, None),
            'cuLaunchKernel(%s)' % name)


_modules = {}


def module(name):
    if name not in _modules:
        _modules[name] = Module(name)
    return _modules[name]


class Timer(object):
    def __init__(self):
        self.e0, self.e1 = C.c_void_p(), C.c_void_p()
        _ck(cu().cuEventCreate(C.byref(self.e0), 0), 'cuEventCreate')
        _ck(cu().cuEventCreate(C.byref(self.e1), 0), 'cuEventCreate')

    def start(self):
        _ck(cu().cuEventRecord(self.e0, None), 'cuEventRecord')

    def stop(self):
        _ck(cu().cuEventRecord(self.e1, None), 'cuEventRecord')
        _ck(cu().cuEventSynchronize(self.e1), 'cuEventSynchronize')
        ms = C.c_float()
        _ck(cu().cuEventElapsedTime(C.byref(ms), self.e0, self.e1), 'cuEventElapsedTime')
        return ms.value


def chunk_iterator(nelements, nthreads_per_block=64, max_blocks=1024):
    first = 0
    while first < nelements:
        left = nelements - first
        blocks = int(left // nthreads_per_block)
        if left % nthreads_per_block != 0:
            blocks += 1
        blocks = min(max_blocks, blocks)
        n = min(left, blocks * nthreads_per_block)
        yield (first, n, blocks)
        first += n


SIZEOF_CURANDSTATE = 48


class RefRNG(object):
    """get_rng_states(size, seed): init_rng<<<size//64+1, 64>>>(size, s, seed, 0)."""

    def __init__(self, size, seed=1):
        self.size = int(size)
        self.mem = DevMem(self.size * SIZEOF_CURANDSTATE)
        module('propagate.cubin').launch('init_rng', self.size // 64 + 1, 64, C.c_int(self.size), self.mem,
                                         C.c_ulonglong(seed), C.c_ulonglong(0))
        sync()

    def states6(self, first=0, count=None):
        """{d, v0..v4} words of each curandStateXORWOW (d at offset 0, v at 4..24)."""
        count = self.size - first if count is None else count
        raw = from_dev(self.mem, np.uint32, (count, 12), offset=first * 48, nbytes=count * 48)
        return np.ascontiguousarray(raw[:, :6])


def _struct(parts):
    """Pack (ctype value) members like make_gpu_struct: pointers 8-aligned, scalars packed."""
    buf = bytearray()
    for p in parts:
        b = bytes(p) if not isinstance(p, (bytes, bytearray)) else p
        if len(b) == 8 and len(buf) % 8:
            raise RefError('cannot align 64-bit pointer')
        buf += b
    return buf


def _ptr(v):
    return np.uint64(v).tobytes()


class RefGeometry(object):
    """Device structs exactly as GPUGeometry builds them, from the same desc."""

    def __init__(self, desc, keep):
        self.keep = keep
        self.vertices = to_dev(keep['vertices'])
        self.triangles = to_dev(keep['triangles'])
        self.codes = to_dev(keep['codes'])
        self.colors = to_dev(keep['colors'])
        self.solid_id = to_dev(keep['solid_id'])
        self.nodes = to_dev(keep['nodes'])
        self.extra_nodes = DevMem(16)
        pool = np.concatenate([keep['pool'], np.zeros(8, np.float32)])
        self.pool = to_dev(pool)
        P = self.pool.ptr
        W, T = desc.wavelength_n, desc.time_n
        self._aux = []
        mat_ptrs = []
        for i in range(desc.nmaterials):
            m = keep['mats'][i]
            comp_ptr = [0, 0, 0, 0]
            if m.num_comp:
                for j, (off, stride) in enumerate(((m.comp_reemission_prob, W), (m.comp_reemission_wvl_cdf, W),
                                                   (m.comp_reemission_time_cdf, T), (m.comp_absorption_length, W))):
                    arr = np.array([P + 4 * (off + c * stride) for c in range(m.num_comp)], dtype=np.uint64)
                    d = to_dev(arr)
                    self._aux.append(d)
                    comp_ptr[j] = d.ptr
            s = _struct([_ptr(P + 4 * m.refractive_index), _ptr(P + 4 * m.absorption_length),
                         _ptr(P + 4 * m.scattering_length), _ptr(comp_ptr[0]), _ptr(comp_ptr[1]), _ptr(comp_ptr[2]),
                         _ptr(comp_ptr[3]), np.uint32(m.num_comp).tobytes(), np.uint32(W).tobytes(),
                         np.float32(desc.wavelength_step).tobytes(), np.float32(desc.wavelength_start).tobytes(),
                         np.uint32(T).tobytes(), np.float32(desc.time_step).tobytes(),
                         np.float32(desc.time_start).tobytes()])
            s += b'\0' * (88 - len(s))
            d = to_dev(np.frombuffer(bytes(s), dtype=np.uint8))
            self._aux.append(d)
            mat_ptrs.append(d.ptr)
        self.material_ptrs = to_dev(np.array(mat_ptrs, dtype=np.uint64))
        surf_ptrs = []
        for i in range(desc.nsurfaces):
            sf = keep['surfs'][i]
            if sf.model < 0:
                surf_ptrs.append(0)
                continue
            dich = 0
            if sf.dichroic_nangles > 0:
                n = sf.dichroic_nangles
                r = to_dev(np.array([P + 4 * (sf.dichroic_reflect + a * W) for a in range(n)], dtype=np.uint64))
                t = to_dev(np.array([P + 4 * (sf.dichroic_transmit + a * W) for a in range(n)], dtype=np.uint64))
                ds = _struct([_ptr(P + 4 * sf.dichroic_angles), _ptr(r.ptr), _ptr(t.ptr), np.uint32(n).tobytes()])
                ds += b'\0' * (32 - len(ds))
                dd = to_dev(np.frombuffer(bytes(ds), dtype=np.uint8))
                self._aux += [r, t, dd]
                dich = dd.ptr
            ang = 0
            if sf.angular_nangles > 0:
                as_ = _struct([_ptr(P + 4 * sf.angular_angles), _ptr(P + 4 * sf.angular_transmit),
                               _ptr(P + 4 * sf.angular_reflect_specular), _ptr(P + 4 * sf.angular_reflect_diffuse),
                               np.uint32(sf.angular_nangles).tobytes()])
                as_ += b'\0' * (40 - len(as_))
                ad = to_dev(np.frombuffer(bytes(as_), dtype=np.uint8))
                self._aux.append(ad)
                ang = ad.ptr
            s = _struct([_ptr(P + 4 * sf.detect), _ptr(P + 4 * sf.absorb), _ptr(P + 4 * sf.reemit),
                         _ptr(P + 4 * sf.reflect_diffuse), _ptr(P + 4 * sf.reflect_specular), _ptr(P + 4 * sf.eta),
                         _ptr(P + 4 * sf.k), _ptr(P + 4 * sf.reemission_cdf), _ptr(dich), _ptr(ang),
                         np.uint32(sf.model).tobytes(), np.uint32(W).tobytes(), np.uint32(sf.transmissive).tobytes(),
                         np.float32(desc.wavelength_step).tobytes(), np.float32(desc.wavelength_start).tobytes(),
                         np.float32(sf.thickness).tobytes()])
            s += b'\0' * (104 - len(s))
            d = to_dev(np.frombuffer(bytes(s), dtype=np.uint8))
            self._aux.append(d)
            surf_ptrs.append(d.ptr)
        self.surface_ptrs = to_dev(np.array(surf_ptrs if surf_ptrs else [0], dtype=np.uint64))
        # analytic wire planes: one 88-byte struct WirePlane each (geometry_types.h:42-58) behind a
        # pointer array, as gpu/geometry.py:343-387 builds them
        wp_ptrs = []
        for i in range(desc.nwireplanes):
            wp = keep['wireplanes'][i]
            ws = _struct([np.array(list(wp.origin) + list(wp.u) + list(wp.v), dtype=np.float32).tobytes(),
                          np.array([wp.pitch, wp.radius, wp.umin, wp.umax, wp.vmin, wp.vmax, wp.v0], dtype=np.float32).tobytes(),
                          np.array([wp.surface_index, wp.material_outer_index, wp.material_inner_index], dtype=np.int32).tobytes(),
                          np.uint32(wp.color).tobytes()])
            assert len(ws) == 80, len(ws)
            dws = to_dev(np.frombuffer(bytes(ws), dtype=np.uint8))
            self._aux.append(dws)
            wp_ptrs.append(dws.ptr)
        self.wireplane_ptrs = to_dev(np.array(wp_ptrs, dtype=np.uint64)) if wp_ptrs else None
        g = _struct([_ptr(self.vertices.ptr), _ptr(self.triangles.ptr), _ptr(self.codes.ptr), _ptr(self.colors.ptr),
                     _ptr(self.nodes.ptr), _ptr(self.extra_nodes.ptr), _ptr(self.material_ptrs.ptr),
                     _ptr(self.surface_ptrs.ptr), _ptr(self.wireplane_ptrs.ptr if wp_ptrs else 0),
                     np.array(list(desc.world_origin), dtype=np.float32).tobytes(),
                     np.float32(desc.world_scale).tobytes(), np.int32(desc.nnodes).tobytes(),
                     np.int32(len(wp_ptrs)).tobytes()])
        assert len(g) == 96, len(g)
        self.gpudata = to_dev(np.frombuffer(bytes(g), dtype=np.uint8))

    def attach_detector(self, detector):
        self.s2c = to_dev(np.asarray(detector.solid_id_to_channel_index, dtype=np.int32))
        self.tx = to_dev(np.asarray(detector.time_cdf[0], dtype=np.float32))
        self.ty = to_dev(np.asarray(detector.time_cdf[1], dtype=np.float32))
        self.qx = to_dev(np.asarray(detector.charge_cdf[0], dtype=np.float32))
        self.qy = to_dev(np.asarray(detector.charge_cdf[1], dtype=np.float32))
        self.nchannels = detector.num_channels()
        s = _struct([_ptr(self.s2c.ptr), _ptr(self.tx.ptr), _ptr(self.ty.ptr), _ptr(self.qx.ptr), _ptr(self.qy.ptr),
                     np.int32(self.nchannels).tobytes(), np.int32(len(detector.time_cdf[0])).tobytes(),
                     np.int32(len(detector.charge_cdf[0])).tobytes(),
                     np.float32(detector.charge_cdf[0][-1] / 2 ** 16).tobytes()])
        assert len(s) == 56
        self.detector_gpu = to_dev(np.frombuffer(bytes(s), dtype=np.uint8))


class RefPhotons(object):
    FIELDS = ('pos', 'dir', 'pol', 'wavelengths', 't', 'last_hit_triangles', 'flags', 'weights', 'evidx')

    def __init__(self, photons):
        self.n = len(photons.pos)
        self.pos = to_dev(np.asarray(photons.pos, dtype=np.float32))
        self.dir = to_dev(np.asarray(photons.dir, dtype=np.float32))
        self.pol = to_dev(np.asarray(photons.pol, dtype=np.float32))
        self.wavelengths = to_dev(np.asarray(photons.wavelengths, dtype=np.float32))
        self.t = to_dev(np.asarray(photons.t, dtype=np.float32))
        self.last_hit_triangles = to_dev(np.asarray(photons.last_hit_triangles, dtype=np.int32))
        self.flags = to_dev(np.asarray(photons.flags, dtype=np.uint32))
        self.weights = to_dev(np.asarray(photons.weights, dtype=np.float32))
        self.evidx = to_dev(np.asarray(photons.evidx, dtype=np.uint32))

    def get(self):
        from chroma_lite_b200 import event
        return event.Photons(*[from_dev(getattr(self, f)) for f in self.FIELDS])

    def get_flat_hits(self, geom, target_flag=0x4, nthreads_per_block=256, max_blocks=1024):
        """GPUPhotons.get_flat_hits (gpu/photon.py:141-209): count kernel -> D2H count ->
        allocate -> compaction kernel -> 10 D2H copies.  Returns dict of host arrays."""
        mod = module('propagate.cubin')
        sync()
        counter = to_dev(np.zeros(1, dtype=np.uint32))
        for first, n, blocks in chunk_iterator(self.n, nthreads_per_block, max_blocks):
            mod.launch('count_photon_hits', blocks, nthreads_per_block, C.c_int(first), C.c_int(n), C.c_uint(target_flag),
                       self.flags, geom.solid_id, self.last_hit_triangles, geom.detector_gpu, counter)
        sync()
        nhit = int(from_dev(counter)[0])
        out = {f: DevMem(max(nhit, 1) * (12 if f in ('pos', 'dir', 'pol') else 4)) for f in self.FIELDS}
        channels = DevMem(max(nhit, 1) * 4)
        if nhit:
            _ck(cu().cuMemsetD32_v2(C.c_uint64(counter.ptr), 0, C.c_size_t(1)), 'memset')
            for first, n, blocks in chunk_iterator(self.n, nthreads_per_block, max_blocks):
                mod.launch('copy_photon_hits', blocks, nthreads_per_block, C.c_int(first), C.c_int(n), C.c_uint(target_flag),
                           geom.solid_id, geom.detector_gpu, counter,
                           self.pos, self.dir, self.wavelengths, self.pol, self.t, self.flags, self.last_hit_triangles,
                           self.weights, self.evidx,
                           out['pos'], out['dir'], out['wavelengths'], out['pol'], out['t'], out['flags'],
                           out['last_hit_triangles'], out['weights'], out['evidx'], channels)
        host = {}
        for f in self.FIELDS:
            if f in ('pos', 'dir', 'pol'):
                host[f] = from_dev(out[f], np.float32, (nhit, 3))
            else:
                host[f] = from_dev(out[f], getattr(self, f).dtype, (nhit,))
        host['channel'] = from_dev(channels, np.int32, (nhit,))
        return host

    def propagate(self, geom, rng, nthreads_per_block=256, max_blocks=1024, max_steps=10, use_weights=False,
                  scatter_first=0, force_single_launch=False):
        """Launch-for-launch replay of GPUPhotons.propagate (gpu/photon.py:240-290).
        force_single_launch=True runs the kernel once with nsteps=max_steps (the
        deterministic replay mode of SURVEY App. A-2; needs rng.size >= n).
        Returns dict(launches, ms)."""
        mod = module('propagate.cubin')
        nphotons = self.n
        iq = np.empty(nphotons + 1, dtype=np.uint32)
        iq[0] = 0
        iq[1:] = np.arange(nphotons, dtype=np.uint32)
        oq = np.zeros(nphotons + 1, dtype=np.uint32)
        oq[0] = 1
        input_queue, output_queue = to_dev(iq), to_dev(oq)
        one = np.ones(1, dtype=np.uint32)
        step, launches = 0, 0
        timer = Timer()
        timer.start()
        while step < max_steps:
            if force_single_launch or nphotons < nthreads_per_block * 16 * 8 or use_weights:
                nsteps = max_steps - step
            else:
                nsteps = 1
            if force_single_launch:
                chunks = [(0, nphotons, (nphotons + nthreads_per_block - 1) // nthreads_per_block)]
            else:
                chunks = chunk_iterator(nphotons, nthreads_per_block, max_blocks)
            for first, n, blocks in chunks:
                mod.launch('propagate', blocks, nthreads_per_block, C.c_int(first), C.c_int(n), input_queue.ptr + 4,
                           output_queue, rng.mem, self.pos, self.dir, self.wavelengths, self.pol, self.t, self.flags,
                           self.last_hit_triangles, self.weights, self.evidx, C.c_int(nsteps), C.c_int(int(use_weights)),
                           C.c_int(scatter_first), geom.gpudata)
                launches += 1
            step += nsteps
            scatter_first = 0
            if step < max_steps:
                input_queue, output_queue = output_queue, input_queue
                _ck(cu().cuMemcpyHtoD_v2(C.c_uint64(output_queue.ptr), one.ctypes.data_as(C.c_void_p), C.c_size_t(4)), 'HtoD')
                nphotons = int(from_dev(input_queue, np.uint32, (1,), nbytes=4)[0]) - 1
                if nphotons == 0:
                    break
        sync()
        return dict(launches=launches, ms=timer.stop())


def intersect(geom, origins, directions, last_hit=None, block=64):
    """ref_intersect: the reference's intersect_mesh on every ray (tri, dist)."""
    o = to_dev(np.asarray(origins, dtype=np.float32))
    d = to_dev(np.asarray(directions, dtype=np.float32))
    n = len(origins)
    lh = to_dev(np.asarray(last_hit, dtype=np.int32)) if last_hit is not None else None
    tri = to_dev(np.full(n, -1, dtype=np.int32))
    dist = to_dev(np.zeros(n, dtype=np.float32))
    timer = Timer()
    timer.start()
    module('ref_wrap.cubin').launch('ref_intersect', n // block + 1, block, C.c_int(n), o, d, lh, geom.gpudata, tri, dist)
    ms = timer.stop()
    return from_dev(tri), from_dev(dist), ms


class ResidentRays(object):
    """Rays and result arrays kept on the device: the reference's intersect_mesh timed launch after launch
    without the upload in between (bench.py --workload rays --impl reference)."""
    def __init__(self, origins, directions):
        self.n = len(origins)
        self.o = to_dev(np.asarray(origins, dtype=np.float32))
        self.d = to_dev(np.asarray(directions, dtype=np.float32))
        self.tri = to_dev(np.full(self.n, -1, dtype=np.int32))
        self.dist = to_dev(np.zeros(self.n, dtype=np.float32))
        self.timer = Timer()

    def launch(self, geom, block=64):
        """One launch over all rays; returns its device time in ms."""
        self.timer.start()
        module('ref_wrap.cubin').launch('ref_intersect', self.n // block + 1, block, C.c_int(self.n), self.o, self.d, None,
                                        geom.gpudata, self.tri, self.dist)
        return self.timer.stop()


def rng_words(n, seed, first_stream=0, offset=0, ndraw=4):
    """(words uint32 (n,ndraw), state6 uint32 (n,6)) straight from curand_init/curand."""
    out = to_dev(np.zeros((n, ndraw), dtype=np.uint32))
    st = to_dev(np.zeros((n, 6), dtype=np.uint32))
    module('ref_wrap.cubin').launch('ref_rng_words', n // 64 + 1, 64, C.c_int(n), C.c_ulonglong(seed),
                                    C.c_ulonglong(first_stream), C.c_ulonglong(offset), C.c_int(ndraw), out, st)
    sync()
    return from_dev(out), from_dev(st)


def run_daq(geom, photons, rng, nthreads_per_block=64, max_blocks=1024, start_photon=0, nphotons=None, weight=1.0):
    """GPUDaq.begin_acquire/acquire/end_acquire (gpu/daq.py:55-101), ndaq=1.
    Returns (t float32, q float32, flags uint32, time_int, q_int)."""
    mod = module('daq.cubin')
    nch = geom.nchannels
    nphotons = photons.n - start_photon if nphotons is None else nphotons
    tint = to_dev(np.zeros(nch, dtype=np.uint32))
    qint = to_dev(np.zeros(nch, dtype=np.uint32))
    hist = to_dev(np.zeros(nch, dtype=np.uint32))
    tf = to_dev(np.zeros(nch, dtype=np.float32))
    qf = to_dev(np.zeros(nch, dtype=np.float32))
    mod.launch('reset_earliest_time_int', nch // 64 + 1, 64, C.c_float(1e9), C.c_int(nch), tint)
    for first, n, blocks in chunk_iterator(nphotons, nthreads_per_block, max_blocks):
        mod.launch('run_daq', blocks, nthreads_per_block, rng.mem, C.c_uint(0x4), C.c_int(start_photon + first), C.c_int(n),
                   photons.t, photons.flags, photons.last_hit_triangles, photons.weights, geom.solid_id, geom.detector_gpu,
                   tint, qint, hist, C.c_float(weight))
    mod.launch('convert_sortable_int_to_float', nch // 64 + 1, 64, C.c_int(nch), tint, tf)
    mod.launch('convert_charge_int_to_float', nch // 64 + 1, 64, geom.detector_gpu, qint, qf)
    sync()
    return from_dev(tf), from_dev(qf), from_dev(hist), from_dev(tint), from_dev(qint)


def run_daq_many(geom, photons, rng, ndaq, nthreads_per_block=64, max_blocks=1024, start_photon=0, nphotons=None,
                 weight=1.0):
    """GPUDaq(ndaq > 1): begin_acquire / acquire with run_daq_many, one block per photon
    (gpu/daq.py:55-59, 80-101).  Returns (t, q, flags) of ndaq * nchannels entries."""
    mod = module('daq.cubin')
    nch = geom.nchannels
    count = nch * ndaq
    nphotons = photons.n - start_photon if nphotons is None else nphotons
    tint = to_dev(np.zeros(count, dtype=np.uint32))
    qint = to_dev(np.zeros(count, dtype=np.uint32))
    hist = to_dev(np.zeros(count, dtype=np.uint32))
    tf = to_dev(np.zeros(count, dtype=np.float32))
    qf = to_dev(np.zeros(count, dtype=np.float32))
    mod.launch('reset_earliest_time_int', count // 64 + 1, 64, C.c_float(1e9), C.c_int(count), tint)
    for first, n, blocks in chunk_iterator(nphotons, 1, max_blocks):
        mod.launch('run_daq_many', blocks, nthreads_per_block, rng.mem, C.c_uint(0x4), C.c_int(start_photon + first),
                   C.c_int(n), photons.t, photons.flags, photons.last_hit_triangles, photons.weights, geom.solid_id,
                   geom.detector_gpu, tint, qint, hist, C.c_int(ndaq), C.c_int(nch), C.c_float(weight))
    mod.launch('convert_sortable_int_to_float', count // 64 + 1, 64, C.c_int(count), tint, tf)
    # convert_charge_int_to_float covers nchannels entries per launch (daq.cu:163-173): one call per copy
    sync()
    q = from_dev(qint).astype(np.float32)
    return from_dev(tf), q, from_dev(hist), from_dev(tint), from_dev(qint)


def make_leaves(vertices, triangles, world_origin, world_scale):
    """The reference's make_leaves kernel (chroma/cuda/bvh.cu:148-203, launched as
    in chroma/gpu/bvh.py:66-78).  Returns (leaf_nodes uint32 (T,4), morton uint64 (T,))."""
    v = to_dev(np.asarray(vertices, dtype=np.float32))
    t = to_dev(np.asarray(triangles, dtype=np.uint32))
    n = len(triangles)
    nodes = to_dev(np.zeros((n, 4), dtype=np.uint32))
    codes = to_dev(np.zeros(n, dtype=np.uint64))

    class float3(C.Structure):
        _fields_ = [('x', C.c_float), ('y', C.c_float), ('z', C.c_float)]
    o = float3(*[float(x) for x in world_origin])
    mod = module('bvh.cubin')
    for first, cnt, blocks in chunk_iterator(n, 256, 30000):
        mod.launch('make_leaves', blocks, 256, C.c_uint(first), C.c_uint(cnt), t, v, o, C.c_float(world_scale), nodes, codes)
    sync()
    return from_dev(nodes), from_dev(codes)


def merge_nodes_detailed(nodes, first_child, nchild):
    """The reference's make_parents_detailed kernel launched as merge_nodes_detailed does
    (chroma/cuda/bvh.cu:269-308, chroma/gpu/bvh.py:84-112): one parent per (first_child, nchild)."""
    child = to_dev(np.ascontiguousarray(nodes, dtype=np.uint32))
    fc = to_dev(np.asarray(first_child).astype(np.int32))
    nc = to_dev(np.asarray(nchild).astype(np.int32))
    nparent = len(first_child)
    parents = to_dev(np.zeros((nparent, 4), dtype=np.uint32))
    mod = module('bvh.cubin')
    for first, cnt, blocks in chunk_iterator(nparent, 256, 10000):
        mod.launch('make_parents_detailed', blocks, 256, C.c_uint(first), C.c_uint(cnt), child, parents, fc, nc)
    sync()
    return from_dev(parents)


def concatenate_layers(layers):
    """The reference's copy_and_offset kernel launched as concatenate_layers does
    (chroma/cuda/bvh.cu:364-384, chroma/gpu/bvh.py:239-267): layers root first into one node array,
    child ids of every layer but the leaves shifted by the end of that layer."""
    bounds = np.insert(np.cumsum([len(l) for l in layers]), 0, 0)
    nodes = to_dev(np.zeros((int(bounds[-1]), 4), dtype=np.uint32))
    mod = module('bvh.cubin')
    for start, end, layer in zip(bounds[:-1], bounds[1:], layers):
        offset = 0 if end == bounds[-1] else int(end)
        src = to_dev(np.ascontiguousarray(layer, dtype=np.uint32))
        for first, cnt, blocks in chunk_iterator(int(end - start), 256, 10000):
            mod.launch('copy_and_offset', blocks, 256, C.c_uint(first), C.c_uint(cnt), C.c_uint(offset), src,
                       nodes.ptr + 16 * int(start))
    sync()
    return from_dev(nodes), bounds


def collapse_chains(nodes, layer_bounds):
    """The reference's collapse_child kernel launched as collapse_chains does
    (chroma/cuda/bvh.cu:530-543, chroma/gpu/bvh.py:114-130): layers bottom up, leaves excluded."""
    dev = to_dev(np.ascontiguousarray(nodes, dtype=np.uint32))
    mod = module('bvh.cubin')
    bounds = list(zip(layer_bounds[:-1], layer_bounds[1:]))[:-1]
    bounds.reverse()
    for start, end in bounds:
        mod.launch('collapse_child', 120, 256, C.c_uint(int(start)), C.c_uint(int(end)), dev)
    sync()
    return from_dev(dev)


# ---------------------------------------------------------------- PDF accumulators
class RefKernelPDF(object):
    """GPUKernelPDF's device side (gpu/pdf.py:44-61, 140-160) on the reference's pdf.cu kernels:
    same launch shapes (block 64, grid n//64+1).  Arrays are host float32/uint32 in, state on the
    device; `channels_t` / `channels_q` are the DAQ output of one acquisition."""

    def __init__(self, nchannels, trange, qrange, time_only=True):
        self.mod = module('pdf.cubin')
        self.n, self.trange, self.qrange, self.time_only = nchannels, trange, qrange, time_only
        z = lambda dt: to_dev(np.zeros(nchannels, dtype=dt))
        self.hitcount, self.tmom1, self.tmom2, self.qmom1, self.qmom2 = z(np.uint32), z(np.float32), z(np.float32), z(np.float32), z(np.float32)

    def accumulate_moments(self, channels_t, channels_q, block=64):
        t, q = to_dev(np.asarray(channels_t, np.float32)), to_dev(np.asarray(channels_q, np.float32))
        self.mod.launch('accumulate_moments', self.n // block + 1, block, C.c_int(int(self.time_only)), C.c_int(self.n), t, q,
                        C.c_float(self.trange[0]), C.c_float(self.trange[1]), C.c_float(self.qrange[0]), C.c_float(self.qrange[1]),
                        self.hitcount, self.tmom1, self.tmom2, self.qmom1, self.qmom2)
        sync()

    def moments(self):
        return tuple(from_dev(a) for a in (self.hitcount, self.tmom1, self.tmom2, self.qmom1, self.qmom2))

    def setup_kernel(self, event_hit, event_time, event_charge, inv_time_bw, inv_charge_bw):
        self.event_hit = to_dev(np.asarray(event_hit, np.uint32))
        self.event_time, self.event_charge = to_dev(np.asarray(event_time, np.float32)), to_dev(np.asarray(event_charge, np.float32))
        self.inv_t, self.inv_q = to_dev(np.asarray(inv_time_bw, np.float32)), to_dev(np.asarray(inv_charge_bw, np.float32))
        self.hitcount = to_dev(np.zeros(self.n, dtype=np.uint32))
        self.tval, self.qval = to_dev(np.zeros(self.n, dtype=np.float32)), to_dev(np.zeros(self.n, dtype=np.float32))

    def accumulate_kernel(self, channels_t, channels_q, block=64):
        t, q = to_dev(np.asarray(channels_t, np.float32)), to_dev(np.asarray(channels_q, np.float32))
        self.mod.launch('accumulate_kernel_eval', self.n // block + 1, block, C.c_int(int(self.time_only)), C.c_int(self.n),
                        self.event_hit, self.event_time, self.event_charge, t, q,
                        C.c_float(self.trange[0]), C.c_float(self.trange[1]), C.c_float(self.qrange[0]), C.c_float(self.qrange[1]),
                        self.inv_t, self.inv_q, self.hitcount, self.tval, self.qval)
        sync()

    def kernel_state(self):
        return from_dev(self.hitcount), from_dev(self.tval), from_dev(self.qval)


class RefPDF(object):
    """GPUPDF's device side (gpu/pdf.py:201-217, 297-330) on the reference's pdf.cu kernels."""

    def __init__(self):
        self.mod = module('pdf.cubin')

    def setup_pdf(self, nchannels, tbins, trange, qbins, qrange):
        self.n, self.tbins, self.trange, self.qbins, self.qrange = nchannels, tbins, trange, qbins, qrange
        self.hitcount = to_dev(np.zeros(nchannels, dtype=np.uint32))
        self.pdf = to_dev(np.zeros((nchannels, tbins, qbins), dtype=np.uint32))

    def add_hits_to_pdf(self, channels_t, channels_q, block=64):
        t, q = to_dev(np.asarray(channels_t, np.float32)), to_dev(np.asarray(channels_q, np.float32))
        self.mod.launch('bin_hits', len(channels_t) // block + 1, block, C.c_int(self.n), q, t, self.hitcount, C.c_int(self.tbins),
                        C.c_float(self.trange[0]), C.c_float(self.trange[1]), C.c_int(self.qbins), C.c_float(self.qrange[0]),
                        C.c_float(self.qrange[1]), self.pdf)
        sync()

    def get_pdfs(self):
        return from_dev(self.hitcount), from_dev(self.pdf)

    def setup_pdf_eval(self, event_hit, event_time, min_twidth, trange, min_bin_content=10):
        event_hit = np.asarray(event_hit)
        self.n = len(event_hit)
        self.nhit = int(np.count_nonzero(event_hit))
        self.hit_to_channel = to_dev(np.flatnonzero(event_hit).astype(np.uint32))
        self.channel_to_hit = to_dev(np.maximum(0, event_hit.astype(np.int64).cumsum() - 1).astype(np.uint32))
        self.event_hit, self.event_time = to_dev(event_hit.astype(np.uint32)), to_dev(np.asarray(event_time, np.float32))
        self.eval_hitcount, self.eval_bincount = to_dev(np.zeros(self.n, np.uint32)), to_dev(np.zeros(self.n, np.uint32))
        self.nearest = to_dev(np.full(max(self.nhit * min_bin_content, 1), 1e9, dtype=np.float32))
        self.min_twidth, self.trange, self.m = min_twidth, trange, min_bin_content

    def accumulate_pdf_eval(self, channels_t, ndaq, block=64, timer=None):
        """channels_t: [ndaq * nchannels] DAQ times of one acquisition (host array, or a DevMem that
        is already on the device).  The work-queue array is filled with 1 on every call, as
        gpu/pdf.py:299-300 does.  With `timer` (a Timer) returns the device time in ms of the fill and
        the two kernels (CUDA events on the launching stream), else None."""
        t = channels_t if isinstance(channels_t, DevMem) else to_dev(np.asarray(channels_t, np.float32))
        nq = max(self.nhit * (ndaq + 1), 1)
        if getattr(self, '_queues', None) is None or self._queues.nbytes < 4 * nq:
            self._queues = DevMem(4 * nq)
        queues = self._queues
        if timer is not None:
            timer.start()
        _ck(cu().cuMemsetD32_v2(C.c_uint64(queues.ptr), C.c_uint(1), C.c_size_t(nq)), 'cuMemsetD32')
        self.mod.launch('accumulate_bincount', self.n // block + 1, block, C.c_int(self.n), C.c_int(ndaq), self.event_hit,
                        self.event_time, t, self.eval_hitcount, self.eval_bincount, C.c_float(self.min_twidth),
                        C.c_float(self.trange[0]), C.c_float(self.trange[1]), C.c_int(self.m), self.channel_to_hit, queues)
        sync()
        if self.nhit:
            self.mod.launch('accumulate_nearest_neighbor_block', self.nhit, block, C.c_int(self.nhit), C.c_int(ndaq),
                            self.hit_to_channel, queues, self.event_time, t, self.nearest, C.c_int(self.m))
        ms = timer.stop() if timer is not None else None
        sync()
        return ms

    def eval_state(self):
        return from_dev(self.eval_hitcount), from_dev(self.eval_bincount), from_dev(self.nearest)
