"""Multi-GPU path (BASELINE config 5): photons sharded over ranks, RNG stream ==
global photon index, per-channel DAQ accumulators combined with MIN / SUM / OR.

  * partition invariance: the reduced arrays are BIT-IDENTICAL for 1, 2, 4 and 8 ranks.  The ranks
    are emulated one after the other on one GPU and combined with cb_daq_reduce_local, which runs
    the same pack / unpack kernels as cb_daq_allreduce with the exchange replaced by a local fold;
  * the real exchange (NCCL inside the library) on two GPUs, when two are visible."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import scenes
from chroma_lite_b200 import gpu, event, parallel, sim, _lib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NEV, NPH, SEED = 8, 30000, 11


def event_photons(e):
    return scenes.point_source(NPH, seed=500 + e, wl_range=(300, 600))


def run_rank(det_gpu, rank, world):
    """One rank's share of the run with the low-level classes; returns its GPUDaq (acquisition open)."""
    plan = parallel.EventPlan(NEV, NPH, rank, world)
    rng = gpu.get_rng_states(plan.nphotons, seed=SEED, first_stream=plan.first_stream)
    daq = gpu.GPUDaq(det_gpu)
    daq.begin_acquire()
    ends = {}
    for e in plan.events:
        gp = gpu.GPUPhotons(event_photons(e))
        window = rng.view(*plan.window(e))
        gp.propagate(det_gpu, window, nthreads_per_block=256, max_blocks=-(-NPH // 256), max_steps=100)
        daq.acquire(gp, window, nthreads_per_block=256, max_blocks=-(-NPH // 256))
        ends[e] = gp.get()
    return daq, ends


def reduced(daqs):
    handles = (C.c_uint64 * len(daqs))(*[d.handle for d in daqs])
    _lib.check(_lib.lib().cb_daq_reduce_local(handles, len(daqs)))
    d = daqs[0]
    return {k: getattr(d, k).get() for k in ('earliest_time_int_gpu', 'channel_q_int_gpu', 'channel_history_gpu',
                                             'earliest_time_gpu', 'channel_q_gpu')}


def test_reduced_daq_is_bit_identical_for_1_2_4_8_ranks(gpu_ready):
    det = scenes.ref_tiny_detector()
    g = gpu.GPUDetector(det)
    base, base_ends = None, None
    for world in (1, 2, 4, 8):
        daqs, ends = [], {}
        for r in range(world):
            d, e = run_rank(g, r, world)
            daqs.append(d)
            ends.update(e)
        out = reduced(daqs)
        if base is None:
            base, base_ends = out, ends
            hit = out['earliest_time_gpu'] < 1e8
            assert hit.sum() > 20 and out['channel_q_int_gpu'][hit].min() > 0      # the run does light up channels
            assert (out['channel_history_gpu'][hit] & event.SURFACE_DETECT).all()
            continue
        for k in base:
            assert np.array_equal(out[k], base[k]), (world, k)
        for e in range(NEV):                                                        # and so is every photon
            for f in ('pos', 'dir', 't', 'flags', 'last_hit_triangles'):
                assert np.array_equal(getattr(ends[e], f), getattr(base_ends[e], f)), (world, e, f)


def test_simulation_with_photon_streams_matches_the_low_level_run(gpu_ready):
    """Simulation(rng_first_stream=...) + simulate(run_daq='accumulate') is the same computation."""
    det = scenes.ref_tiny_detector()
    g = gpu.GPUDetector(det)
    d0, _ = run_rank(g, 0, 1)
    want = reduced([d0])
    outs = []
    for rank in range(2):
        plan = parallel.EventPlan(NEV, NPH, rank, 2)
        s = sim.Simulation(det, seed=SEED, nthreads_per_block=256, max_blocks=-(-NPH // 256),
                           rng_first_stream=plan.first_stream, rng_size=plan.nphotons)
        s.gpu_daq.begin_acquire()
        evs = list(s.simulate((event.Event(photons_beg=event_photons(e)) for e in plan.events), keep_flat_hits=True,
                              keep_hits=False, run_daq='accumulate', max_steps=100, photons_per_batch=NPH))
        assert len(evs) == len(plan.events) and all(len(ev.flat_hits) > 0 for ev in evs)
        outs.append(s)
    got = reduced([s.gpu_daq for s in outs])
    for k in want:
        assert np.array_equal(got[k], want[k]), k


def test_rng_stream_base_and_views(gpu_ready):
    from oracle import orc
    pool = gpu.get_rng_states(1000, seed=3, first_stream=123456789)
    assert np.array_equal(pool.get(0, 16), orc.rng_init(3, 123456789, 16))          # == curand_init(seed, base + i, 0)
    v = pool.view(700, 100)
    assert len(v) == 100 and v.first_stream == 123456789 + 700
    assert np.array_equal(v.get(), pool.get(700, 100))
    with pytest.raises(_lib.ChromaB200Error):
        pool.view(950, 100)


WORKER = r'''
import os, sys, numpy as np
sys.path[:0] = [%(root)r, os.path.join(%(root)r, 'tests')]
import torch, torch.distributed as dist
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
from chroma_lite_b200 import _lib, gpu, parallel
_lib.init(local)
import scenes, test_gpu_multi as T
assert parallel.init_comm() == (rank, world)
det = scenes.ref_tiny_detector()
g = gpu.GPUDetector(det)
mine, _ = T.run_rank(g, rank, world)
mine.allreduce()                               # NCCL inside the library
got = {k: getattr(mine, k).get() for k in ('earliest_time_int_gpu', 'channel_q_int_gpu', 'channel_history_gpu',
                                           'earliest_time_gpu', 'channel_q_gpu')}
whole, _ = T.run_rank(g, 0, 1)                 # the undivided run, on this rank alone
want = T.reduced([whole])
for k in want:
    assert np.array_equal(got[k], want[k]), (rank, k)
parallel.destroy_comm()
dist.barrier()
dist.destroy_process_group()
print('rank %%d of %%d: NCCL-reduced DAQ arrays identical to the single-rank run' %% (rank, world))
'''


def test_nccl_allreduce_inside_the_library(gpu_ready, tmp_path):
    ngpu = _lib.load().cb_device_count()
    if ngpu < 2:
        pytest.skip('needs two GPUs (gpurun --gpus 2); the arithmetic is covered by the one-GPU emulation above')
    world = 2 if ngpu < 4 else 4
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % {'root': ROOT})
    port = str(29600 + os.getpid() % 1000)
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
                          '--master-addr', '127.0.0.1', '--master-port', port, str(script)],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-4000:]
    assert out.stdout.count('identical to the single-rank run') == world, out.stdout[-4000:]
