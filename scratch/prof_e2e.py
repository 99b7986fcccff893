import sys, os, time, cProfile, pstats
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import bench, numpy as np
from chroma_lite_b200 import sim, event, _lib
_lib.init(0)
t={}
det = bench.build_detector('pmt29k', t)
s = sim.Simulation(det, seed=42, nthreads_per_block=512, max_blocks=4883)
ev = bench.make_event(2500000, seed=1000)
kw = dict(keep_hits=False, keep_flat_hits=True, run_daq=True, max_steps=100, photons_per_batch=2500000)
list(s.simulate((event.Event(photons_beg=ev) for _ in range(2)), **kw))
pr = cProfile.Profile(); pr.enable()
t0=time.perf_counter()
list(s.simulate((event.Event(photons_beg=ev) for _ in range(5)), **kw))
print('per event', (time.perf_counter()-t0)/5, s.last_timings)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
