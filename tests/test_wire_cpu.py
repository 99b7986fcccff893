"""RAT bridge frames (bin/chroma-server-rat:30-70): layout and round trips, no GPU."""
import numpy as np
import pytest

from chroma_lite_b200 import event, wire
import scenes


def test_request_layout_and_round_trip():
    ph = scenes.point_source(1000, seed=1, wl_range=(300, 600))
    ph.t[:] = np.linspace(0, 5, len(ph))
    track = np.arange(1000, dtype=np.uint32) * 3
    msg = wire.encode_rat_request(ph, event_id=77, track_id=track)
    assert len(msg) == 8 + 8 * 11 * 1000 + 4 * 1000
    # the reference's own parse of the same bytes (bin/chroma-server-rat:33-35)
    n, evid = np.frombuffer(msg[:8], dtype=np.uint32)
    cols = np.split(np.frombuffer(msg[8:8 + 8 * 11 * n], dtype=np.double), 11)
    assert (n, evid) == (1000, 77)
    assert np.array_equal(cols[0], ph.pos[:, 0].astype(np.float64)) and np.array_equal(cols[9], ph.wavelengths.astype(np.float64))
    got, event_id, tr = wire.decode_rat_request(msg)
    assert event_id == 77 and np.array_equal(tr, track)
    for f in ('pos', 'dir', 'pol', 'wavelengths', 't'):
        assert np.array_equal(getattr(got, f), getattr(ph, f))
    with pytest.raises(ValueError):
        wire.decode_rat_request(msg[:100])
    with pytest.raises(ValueError):
        wire.decode_rat_request(b'\x00')


def test_reply_groups_hits_by_channel_like_the_reference():
    rng = np.random.default_rng(2)
    n = 500
    hits = scenes.point_source(n, seed=3)
    hits.channel = rng.integers(0, 40, n).astype(np.uint32)
    hits.t[:] = np.arange(n)
    msg = wire.encode_rat_reply(hits, 5)
    assert len(msg) == 8 + 4 * 11 * n + 8 * n
    back, evid = wire.decode_rat_reply(msg)
    assert evid == 5
    # the reference builds the reply by concatenating ev.hits[chan] over the channels (:49-56)
    by_chan = {int(c): hits[hits.channel == c] for c in np.unique(hits.channel)}
    expect_t = np.concatenate([by_chan[c].t for c in by_chan])
    expect_c = np.concatenate([np.full(len(by_chan[c]), c, np.uint32) for c in by_chan])
    assert np.array_equal(back.t, expect_t) and np.array_equal(back.channel, expect_c)
    assert np.array_equal(back.pos, np.concatenate([by_chan[c].pos for c in by_chan]))
    # empty reply
    empty, evid = wire.decode_rat_reply(wire.encode_rat_reply(event.Photons(channel=np.zeros(0, np.uint32)), 9))
    assert len(empty) == 0 and evid == 9


def test_photon_server_frames_without_gpu():
    """bin/chroma-server's pickled request/reply around a stand-in simulation (host logic only)."""
    import pickle

    class EchoSim(object):
        def simulate(self, photons, keep_photons_end=False, max_steps=1000, **kw):
            assert keep_photons_end and max_steps == 7
            yield event.Event(photons_end=photons)
    ph = scenes.point_source(50, seed=2)
    ev = pickle.loads(wire.handle_photons_request(EchoSim(), pickle.dumps(ph), max_steps=7))
    assert isinstance(ev, event.Event) and np.array_equal(ev.photons_end.dir, ph.dir)
