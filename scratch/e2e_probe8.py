"""Scratch (torchrun, N ranks): the end-to-end region of bench.py under upload variants, one geometry upload per rank.
Variants: page-locked vs write-combined event buffers x all-zero arrays skipped or sent."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import bench

def main():
    from chroma_lite_b200 import gpu, sim, _lib, parallel, event
    rank, world, local = bench.dist_setup(0)
    _lib.init(local)
    lib = _lib.lib()
    if world > 1:
        parallel.init_comm()
    os.environ.setdefault('CHROMA_B200_TREE_CACHE', bench.cache_dir())
    bench._workload = 'pmt29k'
    tm = {}
    if local == 0:
        det = bench.build_detector('pmt29k', tm)
    bench.barrier(world)
    if local != 0:
        det = bench.build_detector('pmt29k', tm)
    n = 2500000
    s = sim.Simulation(det, seed=42 + rank, cuda_device=local, nthreads_per_block=512, max_blocks=-(-n // 512))
    base = bench.make_event(n, seed=1000 + rank)
    kw = dict(keep_hits=False, keep_flat_hits=True, run_daq=True, max_steps=100, photons_per_batch=n)
    steps = int(os.environ.get('STEPS', '20'))
    run_daq = gpu.GPUDaq(s.gpu_geometry)
    run_daq.begin_acquire(); run_daq.allreduce()
    variants = (('pinned, zeros sent', False, '0'), ('pinned, zeros skipped', False, '1'),
                ('write-combined, zeros sent', True, '0'), ('write-combined, zeros skipped', True, '1'),
                ('pinned, zeros sent', False, '0'), ('pinned, zeros skipped', False, '1'))
    if os.environ.get('VARIANTS'):
        variants = variants[:int(os.environ['VARIANTS'])]
    for name, wc, skip in variants:
        os.environ['CHROMA_B200_SKIP_ZERO'] = skip
        ev = gpu.pin_photons(base, write_combined=wc)
        list(s.simulate((event.Event(photons_beg=ev) for _ in range(5)), **kw))
        bench.barrier(world); lib.cb_synchronize()
        t0 = time.perf_counter()
        run_daq.begin_acquire()
        gaps, tp = [], time.perf_counter()
        for out in s.simulate((event.Event(photons_beg=ev) for _ in range(steps)), **kw):
            gaps.append(time.perf_counter() - tp)
            run_daq.fold(s.gpu_daq, wait=False)
            tp = time.perf_counter()
        run_daq.allreduce().get()
        lib.cb_synchronize(); bench.barrier(world)
        dt = bench.max_over_ranks(time.perf_counter() - t0, world)
        ups = [x[2] - x[1] for x in s.batch_log if x[0] == 'upload' and x[1] >= t0]
        up_max = bench.max_over_ranks(float(np.median(ups)), world)
        first = bench.max_over_ranks(gaps[0], world)
        gap = bench.max_over_ranks(float(np.median(gaps)), world)
        if rank == 0:
            print('   uploads ms:', ' '.join('%.1f' % (u * 1e3) for u in ups), flush=True)
            print('   gpu stages ms (stage, propagate call, kernels, tail):', ' '.join('%.1f/%.1f/%.1f/%.1f' % ((x[2] - x[1]) * 1e3, x[3] * 1e3, x[4], x[5]) for x in s.batch_log if x[0] == 'gpu' and x[1] >= t0), flush=True)
            print('%-32s e2e %.1f M photons/s  (%.1f ms; first yield %.1f ms, median gap %.2f ms, median upload max over ranks %.2f ms, h2d %d B/photon)'
                  % (name, world * n * steps / dt / 1e6, dt * 1e3, first * 1e3, gap * 1e3, up_max * 1e3, s.last_h2d_bytes // n), flush=True)
    if world > 1:
        import torch.distributed as dist
        parallel.destroy_comm(); dist.barrier(); dist.destroy_process_group()

main()
