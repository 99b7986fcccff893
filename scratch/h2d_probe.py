"""Scratch: host->device bandwidth ceiling of this box (pinned, one copy) vs the GPUPhotons upload."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from chroma_lite_b200 import gpu, _lib
from chroma_lite_b200 import gpuarray as ga
import scenes
_lib.init(0)
lib = _lib.lib()
for mb in (30, 130, 512):
    n = mb * 1000000
    h = gpu.pagelocked_empty(n, np.uint8); h[:] = 1
    d = ga.empty(n, np.uint8)
    for _ in range(2): d.set(h)
    t0 = time.perf_counter()
    for _ in range(5): d.set(h)
    dt = (time.perf_counter() - t0) / 5
    print('pinned H2D %4d MB: %.2f ms  %.1f GB/s' % (mb, dt * 1e3, n / dt / 1e9), flush=True)
    hp = np.ones(n, np.uint8)
    t0 = time.perf_counter()
    for _ in range(3): d.set(hp)
    dt = (time.perf_counter() - t0) / 3
    print('pageable H2D %4d MB: %.2f ms  %.1f GB/s' % (mb, dt * 1e3, n / dt / 1e9), flush=True)
    t0 = time.perf_counter()
    for _ in range(5): back = d.get()
    dt = (time.perf_counter() - t0) / 5
    print('D2H (get)  %4d MB: %.2f ms  %.1f GB/s' % (mb, dt * 1e3, n / dt / 1e9), flush=True)
ev = gpu.pin_photons(scenes.point_source(2500000, seed=1, wl_range=(300, 600)))
for _ in range(2): gp = gpu.GPUPhotons(ev, copy_triangles=False, copy_weights=False)
t0 = time.perf_counter()
for _ in range(5): gp = gpu.GPUPhotons(ev, copy_triangles=False, copy_weights=False)
dt = (time.perf_counter() - t0) / 5
print('GPUPhotons(2.5M pinned, 130 MB): %.2f ms  %.1f GB/s' % (dt * 1e3, 130e6 / dt / 1e9))
