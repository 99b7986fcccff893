"""Host logic of Simulation.simulate (no GPU): batching, evidx, the order of the pipeline stages, the
recycling of per-event DAQ objects and what an abandoned generator leaves behind -- with the device
classes replaced by recording stand-ins (chroma/sim.py:225-282 is the behaviour mirrored)."""
import threading
import time
import numpy as np

from chroma_lite_b200 import sim, event


class FakeGpu(object):
    """Stand-ins for chroma_lite_b200.gpu: every device operation becomes an entry in `log`."""

    def __init__(self):
        self.log, self.lock = [], threading.Lock()
        fake = self

        class Marker(object):
            def record(self):
                fake.note('marker.record')
                return self

            def wait(self):
                fake.note('marker.wait')

        class PendingHits(object):
            def __init__(self, bank):
                self.bank = bank

            def get(self, marker=None, ready=False):
                assert ready
                fake.note('hits.get', self.bank.tag)
                n = min(3, len(self.bank.host.pos))
                ph = self.bank.host
                return event.Photons(ph.pos[:n], ph.dir[:n], ph.pol[:n], ph.wavelengths[:n], evidx=self.bank.evidx[:n],
                                     channel=np.arange(n, dtype=np.uint32))

        class GPUPhotons(object):
            count = 0

            def __init__(self, photons, ncopies=1, copy_flags=True, copy_triangles=True, copy_weights=True, evidx_value=None):
                GPUPhotons.count += 1
                self.tag = GPUPhotons.count
                self.host = photons
                self.pos = photons.pos
                self.evidx = (np.full(len(photons.pos), evidx_value, dtype=np.uint32) if evidx_value is not None
                              else np.array(photons.evidx, copy=True))
                fake.note('upload', self.tag, len(photons.pos), evidx_value)
                time.sleep(0.002)

            def propagate(self, *a, **k):
                fake.note('propagate.begin', self.tag)
                time.sleep(0.004)
                fake.note('propagate.end', self.tag)

            def flat_hits_async(self, geometry):
                fake.note('hits.enqueue', self.tag)
                return PendingHits(self)

            def get_flat_hits(self, geometry):
                fake.note('hits.sync', self.tag)
                return PendingHits(self).get(ready=True)

        class Channels(object):
            def __init__(self, daq, tag):
                self.daq, self.tag = daq, tag

            def get(self):
                fake.note('channels.get', self.tag, self.daq.ident)
                return event.Channels(np.zeros(4, bool), np.full(4, float(self.tag), np.float32), np.zeros(4, np.float32),
                                      np.zeros(4, np.uint32))

        class GPUDaq(object):
            made = 0

            def __init__(self, geometry, ndaq=1):
                GPUDaq.made += 1
                self.ident = GPUDaq.made

            def begin_acquire(self):
                fake.note('daq.begin', self.ident)

            def acquire(self, bank, rng, **kw):
                fake.note('daq.acquire', self.ident, bank.tag, kw['start_photon'], kw['nphotons'])
                self.last = bank.tag

            def end_acquire(self):
                return Channels(self, self.last)

            def acquire_async(self, bank, rng, begin=True, finalize=True, **kw):
                fake.note('daq.async', self.ident, bank.tag, kw['start_photon'], kw['nphotons'], begin, finalize)
                return Channels(self, bank.tag)

        self.Marker, self.GPUPhotons, self.GPUDaq = Marker, GPUPhotons, GPUDaq

    def note(self, *what):
        with self.lock:
            self.log.append((threading.current_thread().name,) + what)

    def names(self, *kinds):
        return [x[1:] for x in self.log if x[1] in kinds]


class FakeDetector(object):
    def num_channels(self):
        return 4


def make_sim(monkeypatch):
    fake = FakeGpu()
    monkeypatch.setattr(sim, 'gpu', fake)
    s = sim.Simulation.__new__(sim.Simulation)
    s.detector, s.gpu_geometry = FakeDetector(), object()
    s.nthreads_per_block, s.max_blocks, s.photon_tracking = 64, 1024, False
    s.rng_states, s.rng_per_photon, s.last_timings = object(), False, {}
    s.gpu_daq = fake.GPUDaq(None)
    return s, fake


def photons(n, seed):
    rng = np.random.default_rng(seed)
    p = event.Photons(rng.normal(size=(n, 3)), rng.normal(size=(n, 3)), rng.normal(size=(n, 3)), rng.uniform(300, 600, n))
    p.evidx[:] = 77                      # whatever the caller left there is overwritten (chroma/sim.py:258-260)
    return p


def test_batches_evidx_and_order(monkeypatch):
    s, fake = make_sim(monkeypatch)
    sizes = [50, 60, 200, 10, 10, 10, 300]
    evs = [event.Event(photons_beg=photons(n, k)) for k, n in enumerate(sizes)]
    srcs = [e.photons_beg for e in evs]
    out = list(s.simulate(evs, keep_photons_beg=True, keep_hits=True, keep_flat_hits=True, run_daq=True, photons_per_batch=100))
    assert [e is f for e, f in zip(out, evs)] == [True] * len(evs)               # same objects, same order
    # batches close once they hold >= photons_per_batch photons: [50, 60] [200] [10, 10, 10, 300]
    uploads = fake.names('upload')
    assert [u[2] for u in uploads] == [110, 200, 330]
    assert [u[3] for u in uploads] == [None, 0, None]                            # a one-event batch: evidx filled on the device
    assert [int(p.evidx[0]) for p in srcs] == [0, 1, 0, 0, 1, 2, 3] and all((p.evidx == p.evidx[0]).all() for p in srcs)
    # GPU stages strictly one after the other, each batch: propagate, then hits, then one acquisition per event
    stages = [x for x in fake.log if x[1] in ('propagate.begin', 'propagate.end', 'hits.enqueue', 'daq.async')]
    assert len({x[0] for x in stages}) == 1                                       # all on the one GPU-stage thread
    flat = [(x[1], x[2] if x[1] != 'daq.async' else x[3]) for x in stages]
    expect = []
    for tag, nev in ((1, 2), (2, 1), (3, 4)):
        expect += [('propagate.begin', tag), ('propagate.end', tag), ('hits.enqueue', tag)] + [('daq.async', tag)] * nev
    assert flat == expect
    acq = fake.names('daq.async')
    assert [(a[2], a[3], a[4]) for a in acq] == [(1, 0, 50), (1, 50, 60), (2, 0, 200), (3, 0, 10), (3, 10, 10), (3, 20, 10), (3, 30, 300)]
    assert all(a[5] and a[6] for a in acq)                                        # begin + finalize per event
    # collection happens on the consumer's thread, after the batch's marker, in batch order
    collects = [x for x in fake.log if x[1] in ('marker.wait', 'hits.get')]
    assert {x[0] for x in collects} == {threading.current_thread().name}
    assert [x[2] for x in collects if x[1] == 'hits.get'] == [1, 2, 3]
    # events of a multi-event batch get the hits with their own evidx; channels come from their own acquisition
    assert [len(e.flat_hits) for e in out] == [3, 0, 3, 3, 0, 0, 0]               # the fake returns the batch's first 3 photons
    assert [float(e.channels.t[0]) for e in out] == [1, 1, 2, 3, 3, 3, 3]
    assert all(e.photons_beg is p for e, p in zip(out, srcs))
    # per-event DAQ objects are recycled: never more alive than the pipeline holds at once
    assert fake.GPUDaq.made <= 1 + 4 + 2 + 1


def test_gpu_daq_is_the_yielded_events_and_depth_is_bounded(monkeypatch):
    s, fake = make_sim(monkeypatch)
    evs = [event.Event(photons_beg=photons(100, k)) for k in range(9)]
    seen, in_flight = [], []
    for ev in s.simulate(iter(evs), keep_hits=False, keep_flat_hits=False, run_daq=True, photons_per_batch=100):
        seen.append(s.gpu_daq.ident)
        uploaded = len(fake.names('upload'))
        in_flight.append(uploaded - len(seen))
    daq_of_event = [a[1] for a in fake.names('daq.async')]
    assert seen == daq_of_event                                                   # sim.gpu_daq == the DAQ that acquired this event
    assert max(in_flight) <= s.PIPELINE_DEPTH and len(set(daq_of_event)) <= s.PIPELINE_DEPTH + 1
    assert fake.names('hits.enqueue', 'hits.sync') == []                          # no hits asked for, none compacted


def test_accumulate_mode_and_many_events_per_batch(monkeypatch):
    s, fake = make_sim(monkeypatch)
    run = s.gpu_daq
    out = list(s.simulate([event.Event(photons_beg=photons(40, k)) for k in range(3)], keep_hits=False, keep_flat_hits=True,
                          run_daq='accumulate', photons_per_batch=40))
    acq = fake.names('daq.async')
    assert [a[1] for a in acq] == [run.ident] * 3 and not any(a[5] or a[6] for a in acq)     # one open acquisition, the caller's
    assert all(e.channels is None for e in out) and s.gpu_daq is run
    # more events in a batch than ASYNC_EVENTS_MAX: read back inside the GPU stage, one DAQ for all
    s2, fake2 = make_sim(monkeypatch)
    s2.ASYNC_EVENTS_MAX = 2
    out = list(s2.simulate([event.Event(photons_beg=photons(5, k)) for k in range(5)], keep_hits=False, keep_flat_hits=True,
                           run_daq=True, photons_per_batch=1000))
    assert len(out) == 5 and fake2.names('daq.async') == [] and len(fake2.names('daq.acquire')) == 5
    assert len(fake2.names('hits.sync')) == 1 and fake2.names('marker.record') == []
    assert [int(e.photons_beg is None) for e in out] == [1] * 5                   # keep_photons_beg=False drops them


def test_abandoned_generator_collects_what_is_in_flight(monkeypatch):
    s, fake = make_sim(monkeypatch)
    gen = s.simulate((event.Event(photons_beg=photons(100, k)) for k in range(20)), keep_hits=False, keep_flat_hits=True,
                     run_daq=True, photons_per_batch=100)
    next(gen)
    gen.close()
    started = len(fake.names('propagate.begin'))
    time.sleep(0.05)
    assert len(fake.names('propagate.begin')) == started                          # nothing keeps running behind the caller's back
    assert len(fake.names('propagate.end')) == started == len(fake.names('upload'))
    assert len(fake.names('marker.wait')) >= started                              # every enqueued batch was waited for
