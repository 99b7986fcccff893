"""Scratch: upload time of one 2.5 M-photon event (120 MB, six arrays) from plain and write-combined page-locked
memory, with 1 and 3 upload threads, alone and while propagate calls keep the GPU busy."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import bench
from chroma_lite_b200 import gpu, sim, _lib
import chroma_lite_b200.gpu.photon as gph
_lib.init(0)
bench._workload = 'pmt29k'
det = bench.build_detector('pmt29k', {})
s = sim.Simulation(det, seed=42, cuda_device=0, nthreads_per_block=512, max_blocks=4883)
base = bench.make_event(2500000, seed=1000)
busy = threading.Event(); stop = threading.Event()

def load():
    gp = gpu.GPUPhotons(base); pr = gpu.GPUPhotons(base)
    while not stop.is_set():
        if not busy.is_set():
            time.sleep(0.001); continue
        for f in ('pos', 'dir', 'flags', 'last_hit_triangles', 't'):
            getattr(gp, f).copy_from_device(getattr(pr, f).ptr)
        gp.propagate(s.gpu_geometry, s.rng_states, nthreads_per_block=512, max_blocks=s.max_blocks, max_steps=100)
th = threading.Thread(target=load, daemon=True); th.start()
for wc in (False, True):
    ev = gpu.pin_photons(base, write_combined=wc)
    for threads in (1, 3):
        gph._pool = None
        os.environ['CHROMA_B200_UPLOAD_THREADS'] = str(threads)
        for loaded in (False, True):
            busy.set() if loaded else busy.clear()
            time.sleep(0.05)
            ms = []
            for k in range(12):
                t0 = time.perf_counter()
                gp = gpu.GPUPhotons(ev, copy_triangles=False, copy_weights=False, evidx_value=0)
                ms.append((time.perf_counter() - t0) * 1e3)
            print('write_combined=%s threads=%d gpu_busy=%s: upload ms median %.2f min %.2f' % (wc, threads, loaded, np.median(ms[2:]), min(ms[2:])), flush=True)
stop.set()
