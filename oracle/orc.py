"""ctypes binding of oracle/liborc.so (CPU restatement, chroma_oracle.c).
TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, 'liborc.so')
_lib = None


def build():
    subprocess.check_call(['make', '-s', '-C', _HERE, 'all'])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            build()
        _lib = C.CDLL(_PATH)
        _lib.orc_xorwow_next.restype = C.c_uint32
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def threads():
    """Host threads orc_rng_init / orc_propagate run on (all cores unless set_threads was called)."""
    f = lib().orc_get_threads
    f.restype = C.c_int
    return int(f())


def set_threads(n):
    """0 = all cores.  Results do not depend on the thread count (independent photons / streams)."""
    lib().orc_set_threads(C.c_int(int(n)))


def rng_init(seed, first_stream, n, offset=0):
    """uint32 (n,6) states == curand_init(seed, first_stream+i, offset)."""
    st = np.zeros((n, 6), dtype=np.uint32)
    lib().orc_rng_init(C.c_uint64(seed), C.c_uint64(first_stream), C.c_uint64(n), C.c_uint64(offset), _p(st))
    return st


def rng_words(states, ndraw):
    """Advance each state ndraw times; returns uint32 (n, ndraw)."""
    out = np.zeros((len(states), ndraw), dtype=np.uint32)
    l = lib()
    for i in range(len(states)):
        p = states[i].ctypes.data_as(C.c_void_p)
        for k in range(ndraw):
            out[i, k] = l.orc_xorwow_next(p)
    return out


def rng_fill_uniform(states, low=0.0, high=1.0):
    out = np.zeros(len(states), dtype=np.float32)
    lib().orc_rng_fill_uniform(_p(states), C.c_uint64(len(states)), C.c_float(low), C.c_float(high), _p(out))
    return out


def xorwow_matrix(which, k):
    out = np.zeros(800, dtype=np.uint32)
    lib().orc_xorwow_matrix(C.c_int(which), C.c_int(k), _p(out))
    return out


def intersect(desc, origins, directions, last_hit=None):
    """Reference-order nearest hit for every ray.  Returns (tri, dist, counters)
    with counters = dict(nodes, tris, calls, max_stack)."""
    o = np.ascontiguousarray(origins, dtype=np.float32)
    d = np.ascontiguousarray(directions, dtype=np.float32)
    n = len(o)
    lh = None if last_hit is None else np.ascontiguousarray(last_hit, dtype=np.int32)
    tri = np.full(n, -1, dtype=np.int32)
    dist = np.zeros(n, dtype=np.float32)
    cnt = np.zeros(4, dtype=np.uint64)
    lib().orc_intersect(C.byref(desc), _p(o), _p(d), _p(lh), C.c_uint64(n), _p(tri), _p(dist), _p(cnt))
    return tri, dist, dict(nodes=int(cnt[0]), tris=int(cnt[1]), calls=int(cnt[2]), max_stack=int(cnt[3]))


def triangle_rank(desc):
    rank = np.zeros(desc.ntriangles, dtype=np.uint32)
    lib().orc_triangle_rank(C.byref(desc), _p(rank))
    return rank


class HostBank(object):
    """Host photon arrays in the bank layout (float32 (n,3) etc.)."""
    FIELDS = ('pos', 'dir', 'pol', 'wavelengths', 't', 'last_hit_triangles', 'flags', 'weights', 'evidx')

    def __init__(self, photons):
        self.pos = np.array(photons.pos, dtype=np.float32, order='C')
        self.dir = np.array(photons.dir, dtype=np.float32, order='C')
        self.pol = np.array(photons.pol, dtype=np.float32, order='C')
        self.wavelengths = np.array(photons.wavelengths, dtype=np.float32)
        self.t = np.array(photons.t, dtype=np.float32)
        self.last_hit_triangles = np.array(photons.last_hit_triangles, dtype=np.int32)
        self.flags = np.array(photons.flags, dtype=np.uint32)
        self.weights = np.array(photons.weights, dtype=np.float32)
        self.evidx = np.array(photons.evidx, dtype=np.uint32)

    def struct(self):
        from chroma_lite_b200._lib import CbPhotonBank
        b = CbPhotonBank()
        for f in self.FIELDS:
            setattr(b, f, getattr(self, f).ctypes.data)
        b.n = len(self.pos)
        return b

    def __len__(self):
        return len(self.pos)


def propagate(desc, photons, states, max_steps=10, use_weights=False, scatter_first=0):
    """CPU replay-mode propagate (photon i <-> states[i]).  Returns (HostBank, counters)."""
    bank = photons if isinstance(photons, HostBank) else HostBank(photons)
    assert len(states) >= len(bank)
    b = bank.struct()
    cnt = np.zeros(5, dtype=np.uint64)
    lib().orc_propagate(C.byref(desc), C.byref(b), _p(states), C.c_int(max_steps), C.c_int(int(use_weights)),
                        C.c_int(scatter_first), _p(cnt))
    return bank, dict(nodes=int(cnt[0]), tris=int(cnt[1]), calls=int(cnt[2]), max_stack=int(cnt[3]), steps=int(cnt[4]))


def run_daq(bank, states, detector, solid_map, start=0, n=None, weight=1.0, detection_state=0x4):
    """CPU run_daq (ndaq=1).  Returns (time_int u32, q_int u32, hist u32, charge_unit)."""
    n = len(bank) - start if n is None else n
    nch = detector.num_channels()
    s2c = np.ascontiguousarray(detector.solid_id_to_channel_index, dtype=np.int32)
    tx = np.ascontiguousarray(detector.time_cdf[0], dtype=np.float32)
    ty = np.ascontiguousarray(detector.time_cdf[1], dtype=np.float32)
    qx = np.ascontiguousarray(detector.charge_cdf[0], dtype=np.float32)
    qy = np.ascontiguousarray(detector.charge_cdf[1], dtype=np.float32)
    unit = np.float32(detector.charge_cdf[0][-1] / 2 ** 16)
    tint = np.full(nch, np.float32(1e9).view(np.uint32), dtype=np.uint32)
    qint = np.zeros(nch, dtype=np.uint32)
    hist = np.zeros(nch, dtype=np.uint32)
    sm = np.ascontiguousarray(solid_map, dtype=np.uint32)
    b = bank.struct()
    lib().orc_run_daq(C.byref(b), _p(states), C.c_uint32(detection_state), C.c_uint64(start), C.c_uint64(n), _p(sm),
                      _p(s2c), _p(tx), _p(ty), C.c_int(len(tx)), _p(qx), _p(qy), C.c_int(len(qx)), C.c_float(unit),
                      C.c_float(weight), _p(tint), _p(qint), _p(hist))
    return tint, qint, hist, unit
