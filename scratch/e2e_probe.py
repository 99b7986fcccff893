"""Scratch: per-event timing of Simulation.simulate (e2e pipeline)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import bench
from chroma_lite_b200 import gpu, sim, _lib, event
if len(sys.argv) > 1:
    sys.setswitchinterval(float(sys.argv[1]))
_lib.init(0)
det = bench.build_detector('pmt29k', {})
n = 2500000
s = sim.Simulation(det, seed=42, cuda_device=0, nthreads_per_block=512, max_blocks=max(1024, -(-n // 512)))
ev = gpu.pin_photons(bench.make_event(n, seed=1000))
kw = dict(keep_hits=False, keep_flat_hits=True, run_daq=True, max_steps=100, photons_per_batch=n)
list(s.simulate((event.Event(photons_beg=ev) for _ in range(2)), **kw))
for rep in range(2):
    t0 = time.perf_counter(); ts = []
    for out in s.simulate((event.Event(photons_beg=ev) for _ in range(10)), **kw):
        ts.append(time.perf_counter() - t0)
    print('yield times ms:', ' '.join('%.1f' % (1e3 * t) for t in ts), '| per event %.2f ms' % (1e3 * ts[-1] / 10), s.last_timings, flush=True)
