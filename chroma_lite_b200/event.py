"""Photon / channel / event arrays: the array contract of chroma/event.py.

Field names, dtypes and defaults follow chroma/event.py:3-16, 72-141, 226-309 so
objects are interchangeable with the reference's (duck typing is all the engine
relies on).
"""
import numpy as np

# history bits (chroma/cuda/photon.h:53-68, chroma/event.py:3-16)
NO_HIT = 0x1 << 0
BULK_ABSORB = 0x1 << 1
SURFACE_DETECT = 0x1 << 2
SURFACE_ABSORB = 0x1 << 3
RAYLEIGH_SCATTER = 0x1 << 4
REFLECT_DIFFUSE = 0x1 << 5
REFLECT_SPECULAR = 0x1 << 6
SURFACE_REEMIT = 0x1 << 7
SURFACE_TRANSMIT = 0x1 << 8
BULK_REEMIT = 0x1 << 9
CHERENKOV = 0x1 << 10
SCINTILLATION = 0x1 << 11
# The reference's Python constant is 1<<31 (event.py:16) but its kernel sets bit
# 15 (photon.h:67) and truncates flags to 16 bits; both are exported.
NAN_ABORT = 0x1 << 31
NAN_ABORT_KERNEL = 0x1 << 15

TERMINAL_MASK = NO_HIT | BULK_ABSORB | SURFACE_DETECT | SURFACE_ABSORB | NAN_ABORT_KERNEL

_FIELDS = ('pos', 'dir', 'pol', 'wavelengths', 't', 'last_hit_triangles', 'flags', 'weights',
           'evidx', 'channel')


class Vertex(object):
    def __init__(self, particle_name, pos, dir, ke, t0=0.0, pol=None, steps=None, children=None,
                 trackid=-1, pdgcode=-1):
        self.particle_name, self.pos, self.dir, self.pol = particle_name, pos, dir, pol
        self.ke, self.t0, self.steps, self.children = ke, t0, steps, children
        self.trackid, self.pdgcode = trackid, pdgcode


class Photons(object):
    def __init__(self, pos=np.empty((0, 3)), dir=np.empty((0, 3)), pol=np.empty((0, 3)),
                 wavelengths=np.empty((0)), t=None, last_hit_triangles=None, flags=None,
                 weights=None, evidx=None, channel=None):
        n = len(pos)
        self.pos = np.asarray(pos, dtype=np.float32)
        self.dir = np.asarray(dir, dtype=np.float32)
        self.pol = np.asarray(pol, dtype=np.float32)
        self.wavelengths = np.asarray(wavelengths, dtype=np.float32)
        self.t = np.zeros(n, dtype=np.float32) if t is None else np.asarray(t, dtype=np.float32)
        self.last_hit_triangles = (np.full(n, -1, dtype=np.int32) if last_hit_triangles is None
                                   else np.asarray(last_hit_triangles, dtype=np.int32))
        self.flags = np.zeros(n, dtype=np.uint32) if flags is None else np.asarray(flags, dtype=np.uint32)
        self.weights = np.ones(n, dtype=np.float32) if weights is None else np.asarray(weights, dtype=np.float32)
        self.evidx = np.zeros(n, dtype=np.uint32) if evidx is None else np.asarray(evidx, dtype=np.uint32)
        self.channel = np.zeros(n, dtype=np.uint32) if channel is None else np.asarray(channel, dtype=np.uint32)

    @staticmethod
    def join(photon_list, concatenate=True):
        op = np.concatenate if concatenate else np.asarray
        return Photons(*[op([getattr(p, f) for p in photon_list]) for f in _FIELDS])

    def __add__(self, other):
        return Photons.join([self, other])

    def __len__(self):
        return len(self.pos)

    def __getitem__(self, key):
        return Photons(*[getattr(self, f)[key] for f in _FIELDS])

    def reduced(self, reduction_factor=1.0):
        n = len(self)
        choice = np.random.permutation(n)[:int(n * reduction_factor)]
        return self[choice]

    def __repr__(self):
        return 'Photons[%d]' % len(self.pos)


class Channels(object):
    def __init__(self, hit, t, q, flags=None, evidx=None):
        self.hit, self.t, self.q, self.flags, self.evidx = hit, t, q, flags, evidx

    def hit_channels(self, return_flags=False):
        if return_flags:
            return self.hit.nonzero()[0], self.t[self.hit], self.q[self.hit], self.flags[self.hit]
        return self.hit.nonzero()[0], self.t[self.hit], self.q[self.hit]


class Event(object):
    def __init__(self, id=0, vertices=None, photons_beg=None, photons_end=None, photon_tracks=None,
                 photon_parent_trackids=None, hits=None, flat_hits=None, channels=None):
        self.id = id
        self.nphotons = None
        if vertices is None:
            self.vertices = []
        else:
            self.vertices = vertices if np.iterable(vertices) else [vertices]
        self.photons_beg = photons_beg
        self.photons_end = photons_end
        self.photon_tracks = photon_tracks
        self.photon_parent_trackids = photon_parent_trackids
        self.hits = hits
        self.flat_hits = flat_hits
        self.channels = channels
