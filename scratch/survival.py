"""Scratch: survival function of photon histories in the 29k-PMT detector (CPU oracle)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import bench
from chroma_lite_b200.gpu.geometry import make_desc
from chroma_lite_b200 import event
from oracle import orc
t = {}
det = bench.build_detector('pmt29k', t)
print('built', t, flush=True)
desc, keep = make_desc(det)
n = 100000
ph = bench.make_event(n, seed=999)
for K in (1, 2, 3, 4, 6, 8, 10, 15, 20, 30, 50, 75, 100):
    st = orc.rng_init(42, 0, n)
    bank, cnt = orc.propagate(desc, ph, st, max_steps=K)
    alive = ((bank.flags & event.TERMINAL_MASK) == 0).sum()
    print('max_steps %3d: alive %6d (%.4f%%)  steps total %d' % (K, alive, 100.0 * alive / n, cnt['steps']), flush=True)
