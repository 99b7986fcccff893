"""The request/reply frames of the reference's RAT bridge (bin/chroma-server-rat:30-70), so an
external producer (RAT-PAC's chroma branch, Geant4) can drive the engine: photons in, detected
photons with their channel out.  SURVEY section 8 f-4.

Request  (little endian):  uint32 nphotons, uint32 event id,
                           11 x float64[nphotons]: x y z dx dy dz polx poly polz wavelength t,
                           uint32[nphotons] track id
Reply:                     uint32 nhits, uint32 event id,
                           11 x float32[nhits] in the same order (the photon arrays are float32),
                           uint32[nhits] channel (stand-in for the track id, as in the reference),
                           uint32[nhits] channel
Hits are grouped by ascending channel, photons of a channel in bank order -- the order in which
the reference concatenates ev.hits[chan] over np.unique(channel).

Only the codec and the request handler live here; the ZeroMQ REP loop is `serve_rat` and needs
pyzmq (not a dependency: raises ImportError without it).
"""
import numpy as np

from . import event

_FIELDS = 11


def encode_rat_request(photons, event_id=0, track_id=None):
    """The frame RAT sends (used by tests and by Python producers)."""
    n = len(photons)
    track = np.zeros(n, np.uint32) if track_id is None else np.asarray(track_id, np.uint32)
    cols = [photons.pos[:, 0], photons.pos[:, 1], photons.pos[:, 2], photons.dir[:, 0], photons.dir[:, 1],
            photons.dir[:, 2], photons.pol[:, 0], photons.pol[:, 1], photons.pol[:, 2], photons.wavelengths, photons.t]
    return (np.asarray([n, event_id], np.uint32).tobytes()
            + b''.join(np.ascontiguousarray(c, dtype=np.float64).tobytes() for c in cols) + track.tobytes())


def decode_rat_request(msg):
    """-> (event.Photons, event id, track ids).  The track ids start behind the eleven float64
    arrays (the reference slices them eight bytes early, bin/chroma-server-rat:36, and then
    ignores them)."""
    if len(msg) < 8:
        raise ValueError('RAT request shorter than its header')
    n, event_id = (int(x) for x in np.frombuffer(msg, np.uint32, 2))
    body = 8 * _FIELDS * n
    if len(msg) < 8 + body:
        raise ValueError('RAT request truncated: %d photons need %d bytes, got %d' % (n, 8 + body, len(msg)))
    cols = np.frombuffer(msg, np.float64, _FIELDS * n, offset=8).reshape(_FIELDS, n)
    rest = (len(msg) - 8 - body) // 4
    track = np.frombuffer(msg, np.uint32, min(rest, n), offset=8 + body).copy()
    photons = event.Photons(cols[0:3].T, cols[3:6].T, cols[6:9].T, cols[9], cols[10])
    return photons, event_id, track


def encode_rat_reply(flat_hits, event_id):
    """Reply frame from flat hits (event.Photons with .channel)."""
    order = np.argsort(flat_hits.channel, kind='stable')        # ascending channel, bank order inside
    chan = np.ascontiguousarray(flat_hits.channel[order], dtype=np.uint32)
    f32 = lambda a: np.ascontiguousarray(a[order], dtype=np.float32).tobytes()
    out = [np.asarray([len(order), event_id], np.uint32).tobytes()]
    for vec in (flat_hits.pos, flat_hits.dir, flat_hits.pol):
        out += [f32(vec[:, 0]), f32(vec[:, 1]), f32(vec[:, 2])]
    out += [f32(flat_hits.wavelengths), f32(flat_hits.t), chan.tobytes(), chan.tobytes()]
    return b''.join(out)


def decode_rat_reply(msg):
    """-> (event.Photons with .channel, event id); what the RAT side reads back."""
    n, event_id = (int(x) for x in np.frombuffer(msg, np.uint32, 2))
    cols = np.frombuffer(msg, np.float32, _FIELDS * n, offset=8).reshape(_FIELDS, n)
    chan = np.frombuffer(msg, np.uint32, n, offset=8 + 4 * _FIELDS * n + 4 * n)
    hits = event.Photons(cols[0:3].T, cols[3:6].T, cols[6:9].T, cols[9], cols[10], channel=chan)
    return hits, event_id


def handle_rat_request(sim, msg, max_steps=1000):
    """One request through `sim` (a Simulation on a Detector): propagate, keep the photons detected
    on a channel, no DAQ (RAT does its own), as bin/chroma-server-rat:44-46."""
    photons, event_id, _ = decode_rat_request(msg)
    if len(photons) == 0:
        return encode_rat_reply(event.Photons(channel=np.zeros(0, np.uint32)), event_id)
    ev = next(sim.simulate(photons, keep_photons_beg=False, keep_photons_end=False, keep_hits=False,
                           keep_flat_hits=True, run_daq=False, max_steps=max_steps))
    return encode_rat_reply(ev.flat_hits, event_id)


def serve_rat(sim, address='ipc:///tmp/ipc_chroma', max_requests=None):
    """ZeroMQ REP loop of bin/chroma-server-rat (needs pyzmq)."""
    import zmq
    socket = zmq.Context.instance().socket(zmq.REP)
    socket.bind(address)
    served = 0
    while max_requests is None or served < max_requests:
        socket.send(handle_rat_request(sim, socket.recv()))
        served += 1


# ---- bin/chroma-server:12-40: pickled Photons in, the propagated event out ----------------------
def handle_photons_request(sim, msg, max_steps=1000):
    """One request of the reference's plain photon server: ``msg`` is what zmq's send_pyobj puts on
    the wire (a pickle of a Photons object), the reply is the pickle of the event simulate() yields
    with photons_end kept (bin/chroma-server:31-40).  Like the reference this unpickles what it is
    sent: bind it to trusted peers only."""
    import pickle
    photons_in = pickle.loads(msg)
    ev = next(sim.simulate(photons_in, keep_photons_end=True, max_steps=max_steps))
    return pickle.dumps(ev, pickle.HIGHEST_PROTOCOL)


def serve_photons(sim, address='tcp://*:5024', max_requests=None):
    """ZeroMQ REP loop of bin/chroma-server (needs pyzmq)."""
    import zmq
    socket = zmq.Context.instance().socket(zmq.REP)
    socket.bind(address)
    served = 0
    while max_requests is None or served < max_requests:
        socket.send(handle_photons_request(sim, socket.recv()))
        served += 1
