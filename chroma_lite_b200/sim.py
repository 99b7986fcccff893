"""Simulation driver with the interface of chroma/sim.py:22-282.

Same constructor and simulate() signature; events are batched to
``photons_per_batch`` photons, uploaded once, propagated in one library call,
and hits / DAQ are extracted per event exactly like _simulate_batch does.
"""
import os
import time
import numpy as np

from . import event, gpu
from . import gpuarray as ga


def pick_seed():
    """Seed from the time and the process id (chroma/sim.py:17-20)."""
    return int(time.time()) ^ (os.getpid() << 16) & 2 ** 32 - 1


def _peek(iterable):
    it = iter(iterable)
    first = next(it)

    def chain():
        yield first
        for x in it:
            yield x
    return first, chain()


class Simulation(object):
    def __init__(self, detector, seed=None, cuda_device=None, photon_tracking=False,
                 nthreads_per_block=512, max_blocks=1024):
        self.detector = detector
        self.nthreads_per_block = nthreads_per_block
        self.max_blocks = max_blocks
        self.photon_tracking = photon_tracking
        self.seed = pick_seed() if seed is None else seed
        np.random.seed(self.seed % (2 ** 32))
        self.context = gpu.create_cuda_context(cuda_device)
        if hasattr(detector, 'num_channels'):
            self.gpu_geometry = gpu.GPUDetector(detector)
            self.gpu_daq = gpu.GPUDaq(self.gpu_geometry)
        else:
            self.gpu_geometry = gpu.GPUGeometry(detector)
        # nthreads_per_block*max_blocks is the RNG pool size, part of the replay
        # contract; it no longer dictates a launch shape
        self.rng_states = gpu.get_rng_states(self.nthreads_per_block * self.max_blocks, seed=self.seed)
        self.last_timings = {}

    def _upload_batch(self, batch_events):
        """Host -> device for one batch (runs on the prefetch thread while the previous
        batch propagates; copies use the library's copy stream)."""
        t0 = time.perf_counter()
        sources = [ev.photons_beg for ev in batch_events]
        bounds = np.cumsum(np.concatenate([[0], [len(s) for s in sources]])).astype(np.int64)
        batch = sources[0] if len(sources) == 1 else event.Photons.join(sources)
        gpu_photons = gpu.GPUPhotons(batch, copy_flags=True, copy_triangles=False, copy_weights=False)
        return gpu_photons, bounds, time.perf_counter() - t0

    def _simulate_batch(self, batch_events, keep_photons_beg=False, keep_photons_end=False, keep_hits=True,
                        keep_flat_hits=True, run_daq=False, max_steps=100, verbose=False, uploaded=None):
        t0 = time.perf_counter()
        gpu_photons, bounds, upload_s = uploaded if uploaded is not None else self._upload_batch(batch_events)
        t1 = time.perf_counter()
        tracking = gpu_photons.propagate(self.gpu_geometry, self.rng_states,
                                         nthreads_per_block=self.nthreads_per_block, max_blocks=self.max_blocks,
                                         max_steps=max_steps, track=self.photon_tracking)
        t2 = time.perf_counter()
        is_detector = hasattr(self.detector, 'num_channels')
        if keep_photons_end:
            batch_photons_end = gpu_photons.get()
        if is_detector and (keep_hits or keep_flat_hits):
            batch_hits = gpu_photons.get_flat_hits(self.gpu_geometry)
        t3 = time.perf_counter()
        self.last_timings = {'upload_s': upload_s, 'propagate_s': t2 - t1, 'readback_s': t3 - t2,
                             'nphotons': int(bounds[-1])}
        if verbose:
            print('GPU copy took %0.2f s, propagate %0.2f s' % (t1 - t0, t2 - t1))

        t_daq = 0.0
        for i, (ev, start, end) in enumerate(zip(batch_events, bounds[:-1], bounds[1:])):
            if not keep_photons_beg:
                ev.photons_beg = None
            if self.photon_tracking:
                step_ids, step_photons = tracking
                tracks = [[] for _ in range(end - start)]
                for ids, photons in zip(step_ids, step_photons):
                    mask = np.logical_and(ids >= start, ids < end)
                    if np.count_nonzero(mask) == 0:
                        break
                    sel = photons[mask]
                    for j, pid in enumerate(ids[mask] - start):
                        tracks[pid].append(sel[j])
                ev.photon_tracks = [event.Photons.join(t, concatenate=False) if len(t) > 0 else event.Photons()
                                    for t in tracks]
            if keep_photons_end:
                ev.photons_end = batch_photons_end[start:end]
            if is_detector and (keep_hits or keep_flat_hits):
                ev_hits = batch_hits[batch_hits.evidx == i]
                if keep_hits:
                    ev.hits = {int(c): ev_hits[ev_hits.channel == c] for c in np.unique(ev_hits.channel)}
                if keep_flat_hits:
                    ev.flat_hits = ev_hits
            if hasattr(self, 'gpu_daq') and run_daq:
                # one acquisition per event (chroma/sim.py:141-152)
                td = time.perf_counter()
                self.gpu_daq.begin_acquire()
                self.gpu_daq.acquire(gpu_photons, self.rng_states, start_photon=int(start),
                                     nphotons=int(end - start), nthreads_per_block=self.nthreads_per_block,
                                     max_blocks=self.max_blocks)
                ev.channels = self.gpu_daq.end_acquire().get()
                t_daq += time.perf_counter() - td
            self.last_timings['daq_s'] = t_daq
            self.last_timings['batch_total_s'] = time.perf_counter() - t0
            yield ev

    def simulate(self, iterable, keep_photons_beg=False, keep_photons_end=False, keep_hits=True,
                 keep_flat_hits=True, run_daq=False, max_steps=1000, photons_per_batch=1000000):
        if isinstance(iterable, event.Photons) or (hasattr(iterable, 'pos') and hasattr(iterable, 'wavelengths')):
            first, iterable = iterable, [iterable]
        else:
            first, iterable = _peek(iterable)
        if isinstance(first, event.Event) or hasattr(first, 'photons_beg'):
            pass
        elif isinstance(first, event.Vertex):
            raise NotImplementedError("Vertex input not supported in Chroma")
        else:
            iterable = (event.Event(photons_beg=x) for x in iterable)

        kw = dict(keep_photons_beg=keep_photons_beg, keep_photons_end=keep_photons_end, keep_hits=keep_hits,
                  keep_flat_hits=keep_flat_hits, run_daq=run_daq, max_steps=max_steps)

        def batches():
            nphotons, batch = 0, []
            for ev in iterable:
                ev.nphotons = len(ev.photons_beg)
                evidx = getattr(ev.photons_beg, 'evidx', None)
                if evidx is not None:
                    if isinstance(evidx, ga.DeviceArray):
                        if ev.nphotons:
                            evidx[:ev.nphotons].fill(np.uint32(len(batch)))
                    else:
                        evidx[:ev.nphotons] = np.uint32(len(batch))
                nphotons += ev.nphotons
                batch.append(ev)
                if nphotons >= photons_per_batch:
                    yield batch
                    nphotons, batch = 0, []
            if batch:
                yield batch

        # Double-buffered pipeline: while batch k propagates (the C call releases the
        # GIL), a worker thread uploads batch k+1 on the copy stream.  The reference
        # does upload -> propagate -> download strictly in sequence (sim.py:79-110).
        import concurrent.futures
        it = batches()
        nxt = next(it, None)
        if nxt is None:
            return
        with concurrent.futures.ThreadPoolExecutor(max_workers=1) as pool:
            pending = pool.submit(self._upload_batch, nxt)
            while nxt is not None:
                cur, uploaded = nxt, pending.result()
                nxt = next(it, None)
                if nxt is not None:
                    pending = pool.submit(self._upload_batch, nxt)
                yield from self._simulate_batch(cur, uploaded=uploaded, **kw)

    def __del__(self):
        try:
            self.context.pop()
        except Exception:
            pass
