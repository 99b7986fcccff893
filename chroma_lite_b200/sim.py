"""Simulation driver with the interface of chroma/sim.py:22-282.

Same constructor and simulate() signature; events are batched to
``photons_per_batch`` photons, uploaded once, propagated in one library call,
and hits / DAQ are extracted per event exactly like _simulate_batch does.
"""
import os
import time
from types import SimpleNamespace
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
                 nthreads_per_block=512, max_blocks=1024, rng_first_stream=None, rng_size=None):
        """Same arguments as chroma/sim.py:22-52, plus the multi-GPU extension: with
        ``rng_first_stream`` (global index of this rank's first photon) the pool holds ``rng_size``
        states, one per photon this Simulation will ever see, and consecutive batches take
        consecutive windows of it: RNG stream == global photon index, results do not depend on how a
        run is partitioned over ranks (parallel.EventPlan)."""
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
        self.rng_per_photon = rng_first_stream is not None
        if self.rng_per_photon:
            self.rng_states = gpu.get_rng_states(int(rng_size), seed=self.seed, first_stream=int(rng_first_stream))
            self.rng_cursor = 0               # photons that have taken their window so far
        else:
            self.rng_states = gpu.get_rng_states(self.nthreads_per_block * self.max_blocks, seed=self.seed)
        # several ranks per host: pipeline threads wait blocking instead of spinning (parallel.py)
        import os
        from . import parallel, _lib
        if not os.environ.get('CHROMA_B200_SYNC') and parallel.host_threads_should_block():
            _lib.check(_lib.lib().cb_set_blocking_sync(1))
        self.last_timings = {}

    def _log(self, stage, t0, t1, *extra):
        """(stage, start, end) of the last pipeline stages, perf_counter seconds: where a run's time went
        (bench.py prints the first batches of the end-to-end region from it)."""
        log = self.__dict__.get('batch_log')
        if log is None:
            import collections
            log = self.batch_log = collections.deque(maxlen=256)
        log.append((stage, t0, t1) + extra)

    def _upload_batch(self, batch_events):
        """Host -> device for one batch (runs on the prefetch thread while the previous
        batch propagates; copies use the library's copy stream)."""
        t0 = time.perf_counter()
        sources = [ev.photons_beg for ev in batch_events]
        bounds = np.cumsum(np.concatenate([[0], [len(s) for s in sources]])).astype(np.int64)
        if len(sources) == 1:
            batch = sources[0]
        else:
            # device-resident sources are concatenated on the device (chroma/sim.py:155-226);
            # anything else goes through the host join like the reference's fallback
            batch = self._stack_gpu_photon_sources(sources, copy_flags=True, copy_triangles=False,
                                                   copy_weights=False)
            if batch is None:
                batch = event.Photons.join(sources)
        # the first large batch of a size: cache the device blocks of every batch the pipeline can hold in flight
        # now, while nothing runs (a cudaMalloc met later waits for the kernels in flight)
        n = int(bounds[-1])
        reserve = getattr(gpu, 'reserve_banks', None)
        if reserve is not None and n >= 200000 and n not in self.__dict__.setdefault('_reserved', set()):
            self._reserved.add(n)
            reserve(n, self.PIPELINE_DEPTH + 2)
        # a batch of one event: its evidx is 0 throughout (simulate() has just written that into the host
        # array too), so the device array is filled in place instead of crossing PCIe
        gpu_photons = gpu.GPUPhotons(batch, copy_flags=True, copy_triangles=False, copy_weights=False,
                                     evidx_value=0 if len(sources) == 1 and getattr(batch, 'evidx', None) is not None else None)
        t1 = time.perf_counter()
        self._log('upload', t0, t1)
        self.last_h2d_bytes = getattr(gpu_photons, 'h2d_bytes', None)      # what this batch really sent, when known
        return gpu_photons, bounds, t1 - t0

    @staticmethod
    def _is_gpu_photon_source(photons, copy_flags=True, copy_triangles=False, copy_weights=False):
        """True when every field GPUPhotons would copy is already a device array
        (chroma/sim.py:155-168)."""
        fields = ['pos', 'dir', 'pol', 'wavelengths', 't', 'evidx']
        fields += ['flags'] * bool(copy_flags) + ['last_hit_triangles'] * bool(copy_triangles)
        fields += ['weights'] * bool(copy_weights)
        return all(isinstance(getattr(photons, f, None), ga.DeviceArray) for f in fields)

    @classmethod
    def _stack_gpu_photon_sources(cls, photon_sources, copy_flags=True, copy_triangles=False, copy_weights=False):
        """Concatenate device-resident photon sources into one device-resident source with
        device-to-device copies (chroma/sim.py:170-226); None unless every source qualifies."""
        if not photon_sources or not all(cls._is_gpu_photon_source(p, copy_flags, copy_triangles, copy_weights)
                                         for p in photon_sources):
            return None
        counts = [gpu.photon._resolve_nphotons(p) for p in photon_sources]
        total = int(sum(counts))
        fields = ['pos', 'dir', 'pol', 'wavelengths', 't', 'evidx']
        fields += ['flags'] * bool(copy_flags) + ['last_hit_triangles'] * bool(copy_triangles)
        fields += ['weights'] * bool(copy_weights)
        stacked = {}
        for f in fields:
            dest = ga.empty(total, getattr(photon_sources[0], f).dtype)
            offset = 0
            for p, n in zip(photon_sources, counts):
                if n:
                    dest[offset:offset + n].copy_from_device(getattr(p, f), n * dest.dtype.itemsize)
                offset += n
            stacked[f] = dest
        return SimpleNamespace(true_nphotons=total, **stacked)

    # events per batch up to which the hit read-back and the DAQ of a batch are only ENQUEUED by the GPU stage
    # and collected by the host stage (each event of a batch in flight holds its own per-channel arrays)
    ASYNC_EVENTS_MAX = 8

    def _daq_take(self):
        """A GPUDaq for one event in flight (a small free list: created on demand, reused)."""
        free = self.__dict__.setdefault('_daq_free', [])
        return free.pop() if free else gpu.GPUDaq(self.gpu_geometry)

    def _gpu_stage(self, batch_events, uploaded=None, keep_photons_end=False, keep_hits=True, keep_flat_hits=True,
                   run_daq=False, max_steps=100, verbose=False, defer=False, **_unused):
        """Everything of one batch that needs the GPU, in the reference's order (chroma/sim.py:82-152):
        propagate, photons_end / flat-hit read-back, then one DAQ acquisition per event.  Returns the
        raw host arrays; building the Event objects from them (_host_stage) needs no GPU and overlaps
        with the next batch's GPU stage.

        defer=True (the pipeline of simulate()): hit compaction and DAQ are enqueued behind the propagate
        kernels without a host round trip, a completion marker is recorded, and this thread goes straight
        on to the next batch; _host_stage waits for the marker and reads the results back while the GPU
        already propagates the next batch (the reference: count kernel, 4-byte read-back, copy kernel, ten
        read-backs, DAQ kernels, three read-backs, all with the GPU otherwise idle)."""
        t0 = time.perf_counter()
        if uploaded is not None and hasattr(uploaded, 'result'):
            uploaded = uploaded.result()
        t0b = time.perf_counter()
        gpu_photons, bounds, upload_s = uploaded if uploaded is not None else self._upload_batch(batch_events)
        t1 = time.perf_counter()
        raw = {'bounds': bounds}
        rng, max_blocks = self.rng_states, self.max_blocks
        if self.rng_per_photon:
            # this batch's photons take the next window of the pool: one stream per photon, one chunk
            n = int(bounds[-1])
            rng = self.rng_states.view(self.rng_cursor, n)
            self.rng_cursor += n
            max_blocks = max(max_blocks, -(-n // self.nthreads_per_block))
        raw['tracking'] = gpu_photons.propagate(self.gpu_geometry, rng,
                                                nthreads_per_block=self.nthreads_per_block, max_blocks=max_blocks,
                                                max_steps=max_steps, track=self.photon_tracking)
        t2 = time.perf_counter()
        is_detector = hasattr(self.detector, 'num_channels')
        defer = defer and len(batch_events) <= self.ASYNC_EVENTS_MAX
        if keep_photons_end:
            raw['photons_end'] = gpu_photons.get()
        if is_detector and (keep_hits or keep_flat_hits):
            if defer:
                raw['pending_hits'] = gpu_photons.flat_hits_async(self.gpu_geometry)
            else:
                raw['hits'] = gpu_photons.get_flat_hits(self.gpu_geometry)
        t3 = time.perf_counter()
        if hasattr(self, 'gpu_daq') and run_daq:
            # one acquisition per event (chroma/sim.py:141-152).  run_daq='accumulate' (extension): the
            # events add to ONE acquisition the caller opened with gpu_daq.begin_acquire() and closes with
            # end_acquire() / allreduce() -- the run-level per-channel accumulators of a sharded run
            accumulate = run_daq == 'accumulate'
            if not accumulate:
                raw['pending_channels' if defer else 'channels'] = []
            for start, end in zip(bounds[:-1], bounds[1:]):
                ev_rng = rng.view(int(start), int(end - start)) if self.rng_per_photon and len(bounds) > 2 else rng
                kw = dict(start_photon=int(start), nphotons=int(end - start), nthreads_per_block=self.nthreads_per_block,
                          max_blocks=max_blocks)
                if defer:
                    daq = self.gpu_daq if accumulate else self._daq_take()
                    ch = daq.acquire_async(gpu_photons, ev_rng, begin=not accumulate, finalize=not accumulate, **kw)
                    if not accumulate:
                        raw['pending_channels'].append((daq, ch))
                    continue
                if not accumulate:
                    self.gpu_daq.begin_acquire()
                self.gpu_daq.acquire(gpu_photons, ev_rng, **kw)
                if not accumulate:
                    raw['channels'].append(self.gpu_daq.end_acquire().get())
        if defer:
            markers = self.__dict__.setdefault('_markers_free', [])
            raw['marker'] = (markers.pop() if markers else gpu.Marker()).record()
            raw['bank'] = gpu_photons          # in use by the enqueued kernels until the marker has passed
        t4 = time.perf_counter()
        st = getattr(gpu_photons, 'last_stats', None)
        self._log('gpu', t0b, t4, round(t2 - t1, 6), round(getattr(st, 'kernel_ms', 0.0), 3), round(getattr(st, 'tail_ms', 0.0), 3))
        self.last_timings = {'upload_s': upload_s, 'upload_wait_s': t0b - t0, 'propagate_s': t2 - t1, 'readback_s': t3 - t2,
                             'daq_s': t4 - t3, 'nphotons': int(bounds[-1]), 'batch_total_s': t4 - t0, 'deferred': bool(defer)}
        if verbose:
            print('GPU copy took %0.2f s, propagate %0.2f s' % (t1 - t0, t2 - t1))
        return raw

    def _collect(self, raw):
        """Second half of a deferred GPU stage: wait for the batch's marker, read hits and channels back."""
        marker = raw.pop('marker', None)
        if marker is None:
            return raw
        t0 = time.perf_counter()
        marker.wait()
        if 'pending_hits' in raw:
            raw['hits'] = raw.pop('pending_hits').get(marker, ready=True)
        if 'pending_channels' in raw:
            raw['channels'], raw['daqs'] = [], []
            for daq, ch in raw.pop('pending_channels'):
                raw['channels'].append(ch.get())
                raw['daqs'].append(daq)
        raw.pop('bank', None)
        self._markers_free.append(marker)
        t1 = time.perf_counter()
        self._log('collect', t0, t1)
        self.last_timings = dict(self.last_timings, collect_s=t1 - t0)
        return raw

    def _host_stage(self, batch_events, raw, keep_photons_beg=False, keep_photons_end=False, keep_hits=True,
                    keep_flat_hits=True, **_unused):
        """Slice the batch results back into the events (chroma/sim.py:112-154); host only."""
        bounds = raw['bounds']
        for i, (ev, start, end) in enumerate(zip(batch_events, bounds[:-1], bounds[1:])):
            if not keep_photons_beg:
                ev.photons_beg = None
            if self.photon_tracking:
                step_ids, step_photons = raw['tracking']
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
            if keep_photons_end and 'photons_end' in raw:
                ev.photons_end = raw['photons_end'][start:end]
            if 'hits' in raw:
                batch_hits = raw['hits']
                ev_hits = batch_hits if len(batch_events) == 1 else batch_hits[batch_hits.evidx == i]
                if keep_hits:
                    ev.hits = {int(c): ev_hits[ev_hits.channel == c] for c in np.unique(ev_hits.channel)}
                if keep_flat_hits:
                    ev.flat_hits = ev_hits
            if 'channels' in raw:
                ev.channels = raw['channels'][i]
            if 'daqs' in raw:
                # the event's per-channel arrays on the device: `gpu_daq` is the one of the event being
                # yielded (what a caller folds into run-level accumulators), and goes back to the free list
                # when the consumer comes back for the next event
                self.gpu_daq = raw['daqs'][i]
            yield ev
            if 'daqs' in raw:
                self._daq_free.append(raw['daqs'][i])

    def _simulate_batch(self, batch_events, uploaded=None, **kw):
        raw = self._gpu_stage(batch_events, uploaded=uploaded, **kw)
        yield from self._host_stage(batch_events, raw, **kw)

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

        def set_evidx(ev, index):
            evidx = getattr(ev.photons_beg, 'evidx', None)
            if evidx is None:
                return
            if isinstance(evidx, ga.DeviceArray):
                if ev.nphotons:
                    evidx[:ev.nphotons].fill(np.uint32(index))
            else:
                evidx[:ev.nphotons] = np.uint32(index)

        def batches():
            """Lists of events of >= photons_per_batch photons, each event's evidx set to its place in the
            batch (chroma/sim.py:256-262).  A batch of ONE host-array event is yielded before that write:
            its upload fills evidx on the device (_upload_batch), so the 4 bytes per photon written on the
            host here need not delay it -- the caller does the write once the upload is under way (`late`)."""
            nphotons, batch = 0, []
            for ev in iterable:
                ev.nphotons = len(ev.photons_beg)
                nphotons += ev.nphotons
                batch.append(ev)
                if nphotons >= photons_per_batch:
                    yield self._close_batch(batch, set_evidx)
                    nphotons, batch = 0, []
            if batch:
                yield self._close_batch(batch, set_evidx)

        # Pipeline, up to DEPTH batches in flight: a worker thread uploads batches on its copy stream as
        # soon as they exist, a second one runs their GPU stages strictly one after the other (so the RNG
        # pool is consumed in the reference's order: propagate k, DAQ k, propagate k+1, ...; the C calls
        # release the GIL) and only ENQUEUES hit compaction and DAQ behind the propagate kernels, and this
        # thread waits for a batch's completion marker, reads its results back and turns them into
        # events -- while the GPU already propagates the next batch.  The reference does upload ->
        # propagate -> download strictly in sequence (sim.py:79-110).
        import collections
        it = batches()
        up_pool, gpu_pool = self._workers()
        pending = collections.deque()             # (batch, future of its raw results)
        defer = os.environ.get('CHROMA_B200_DEFER', '1') != '0'      # 0: read every batch back inside its GPU stage

        def submit(item):
            batch, late = item
            up = up_pool.submit(self._upload_batch, batch)
            pending.append((batch, gpu_pool.submit(self._gpu_stage, batch, up, defer=defer, **kw)))
            for ev in late:
                set_evidx(ev, 0)

        try:
            depth = int(os.environ.get('CHROMA_B200_PIPELINE_DEPTH', self.PIPELINE_DEPTH))
            for _ in range(max(depth, 1)):
                nxt = next(it, None)
                if nxt is None:
                    break
                submit(nxt)
            while pending:
                batch, fut = pending.popleft()
                raw = self._collect(fut.result())
                nxt = next(it, None)
                if nxt is not None:
                    submit(nxt)
                yield from self._host_stage(batch, raw, **kw)
        finally:
            # a consumer that stops early must not leave work behind that still uses the RNG pool
            for _, f in pending:
                try:
                    self._collect(f.result())
                except Exception:
                    pass

    PIPELINE_DEPTH = 3

    @staticmethod
    def _close_batch(batch, set_evidx):
        """(batch, events whose evidx is still to be written on the host)."""
        single_host_event = len(batch) == 1 and not isinstance(getattr(batch[0].photons_beg, 'evidx', None), ga.DeviceArray)
        if single_host_event:
            return batch, list(batch)
        for i, ev in enumerate(batch):
            set_evidx(ev, i)
        return batch, []

    # ------------------------------------------------------------------ PDFs / likelihood
    # Upstream Chroma's Simulation.create_pdf / eval_pdf / eval_kernel, which this fork of the
    # reference dropped while keeping gpu/pdf.py and cuda/pdf.cu: Monte
    # Carlo events -> propagate -> DAQ -> per-channel accumulators (gpu.GPUPDF / gpu.GPUKernelPDF).
    def _mc_channels(self, iterable, nreps=1, ndaq=1, max_steps=100):
        """GPUChannels of every (event, repetition, DAQ pass) of the Monte Carlo."""
        assert hasattr(self, 'gpu_daq'), 'PDFs need a detector with channels'
        if isinstance(iterable, event.Photons) or hasattr(iterable, 'wavelengths'):
            iterable = [iterable]
        for item in iterable:
            photons = item.photons_beg if hasattr(item, 'photons_beg') else item
            gpu_photons = gpu.GPUPhotons(photons, ncopies=nreps)
            gpu_photons.propagate(self.gpu_geometry, self.rng_states, nthreads_per_block=self.nthreads_per_block,
                                  max_blocks=self.max_blocks, max_steps=max_steps)
            n = gpu_photons.true_nphotons
            for rep in range(nreps):
                for _ in range(ndaq):
                    self.gpu_daq.begin_acquire()
                    self.gpu_daq.acquire(gpu_photons, self.rng_states, start_photon=rep * n, nphotons=n,
                                         nthreads_per_block=self.nthreads_per_block, max_blocks=self.max_blocks)
                    yield self.gpu_daq.end_acquire()

    def create_pdf(self, iterable, tbins, trange, qbins, qrange, nreps=1, max_steps=100):
        """(hit count per channel, [channel, time, charge] histogram) of the Monte Carlo events."""
        if getattr(self, 'gpu_pdf', None) is None:
            self.gpu_pdf = gpu.GPUPDF()
        config = (tbins, tuple(trange), qbins, tuple(qrange))
        if getattr(self, 'pdf_config', None) != config:
            self.pdf_config = config
            self.gpu_pdf.setup_pdf(self.gpu_geometry.nchannels, tbins, trange, qbins, qrange)
        else:
            self.gpu_pdf.clear_pdf()
        for channels in self._mc_channels(iterable, nreps=nreps, max_steps=max_steps):
            self.gpu_pdf.add_hits_to_pdf(channels)
        return self.gpu_pdf.get_pdfs()

    def eval_pdf(self, event_channels, iterable, min_twidth, trange, min_qwidth, qrange, min_bin_content=100,
                 nreps=1, ndaq=1, time_only=True, max_steps=100):
        """(hit count, PDF value, uncertainty) per channel at the times of `event_channels`
        (event.Channels), estimated from the Monte Carlo events with an adaptive bin."""
        if getattr(self, 'gpu_pdf', None) is None:
            self.gpu_pdf = gpu.GPUPDF()
        self.gpu_pdf.setup_pdf_eval(event_channels.hit, event_channels.t, event_channels.q, min_twidth, trange,
                                    min_qwidth, qrange, min_bin_content=min_bin_content, time_only=time_only)
        for channels in self._mc_channels(iterable, nreps=nreps, ndaq=ndaq, max_steps=max_steps):
            self.gpu_pdf.accumulate_pdf_eval(channels)
        return self.gpu_pdf.get_pdf_eval()

    def eval_kernel(self, event_channels, kernel_generator, trange, qrange, nreps=1, ndaq=1, time_only=True,
                    scale_factor=1.0, max_steps=100):
        """Kernel-density version of eval_pdf: `kernel_generator` is iterated twice (a list, or a
        callable returning a fresh iterable): first for the moments that set the bandwidths, then
        for the kernel sums."""
        source = kernel_generator if callable(kernel_generator) else (lambda: kernel_generator)
        if getattr(self, 'gpu_pdf_kernel', None) is None:
            self.gpu_pdf_kernel = gpu.GPUKernelPDF()
        k = self.gpu_pdf_kernel
        k.setup_moments(self.gpu_geometry.nchannels, trange, qrange, time_only=time_only)
        for channels in self._mc_channels(source(), nreps=nreps, ndaq=ndaq, max_steps=max_steps):
            k.accumulate_moments(channels)
        k.compute_bandwidth(event_channels.hit, event_channels.t, event_channels.q, scale_factor=scale_factor)
        k.setup_kernel(event_channels.hit, event_channels.t, event_channels.q)
        for channels in self._mc_channels(source(), nreps=nreps, ndaq=ndaq, max_steps=max_steps):
            k.accumulate_kernel(channels)
        return k.get_kernel_eval()

    def _workers(self):
        """The two pipeline threads live as long as the Simulation (their CUDA per-thread state --
        device binding, copy stream -- is set up once, not per simulate() call)."""
        if getattr(self, '_pools', None) is None:
            import concurrent.futures
            self._pools = (concurrent.futures.ThreadPoolExecutor(max_workers=1, thread_name_prefix='cb-upload-stage'),
                           concurrent.futures.ThreadPoolExecutor(max_workers=1, thread_name_prefix='cb-gpu-stage'))
        return self._pools

    def __del__(self):
        try:
            for p in (getattr(self, '_pools', None) or ()):
                p.shutdown(wait=False)
        except Exception:
            pass
        try:
            self.context.pop()
        except Exception:
            pass
