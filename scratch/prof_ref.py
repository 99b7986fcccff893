import sys, os, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import bench, numpy as np
from oracle import ref_driver
from chroma_lite_b200 import _lib
from chroma_lite_b200.gpu.geometry import make_desc
_lib.init(0)
t={}
det = bench.build_detector('pmt29k', t)
desc, keep = make_desc(det)
rg = ref_driver.RefGeometry(desc, keep); rg.attach_detector(det)
ev = bench.make_event(2500000, seed=1000)
rng = ref_driver.RefRNG(512*1024, seed=42)
for it in range(3):
    ref_driver.sync(); t0=time.perf_counter()
    rp = ref_driver.RefPhotons(ev); ref_driver.sync(); t1=time.perf_counter()
    r = rp.propagate(rg, rng, nthreads_per_block=512, max_blocks=1024, max_steps=100); t2=time.perf_counter()
    hits = rp.get_flat_hits(rg); ref_driver.sync(); t3=time.perf_counter()
    ch = ref_driver.run_daq(rg, rp, rng, nthreads_per_block=512, max_blocks=1024); t4=time.perf_counter()
    del rp; ref_driver.sync(); t5=time.perf_counter()
    print('upload %.1f propagate %.1f (kernel ms %.1f) hits %.1f daq %.1f free %.1f ms' % ((t1-t0)*1e3,(t2-t1)*1e3,r['ms'],(t3-t2)*1e3,(t4-t3)*1e3,(t5-t4)*1e3), len(hits['t']))
