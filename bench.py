#!/usr/bin/env python
"""bench.py -- photons propagated per second on the 29k-PMT water-Cherenkov
detector (BASELINE.json metric; config 3 of BASELINE.md), 1..8 B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload pmt29k|tiny|rays] [--photons P]

A "step" = one event of P photons (default 2.5 M, isotropic point source at the
origin, lambda ~ U(300,600) nm) propagated to termination (max_steps=100) in the
detector, photons already resident in HBM (`value`), or the same event pushed
through the public API Simulation.simulate from HOST arrays incl. upload, hit
read-back and DAQ (`e2e`).  Photons shard across ranks (weak scaling: every rank
propagates its own P-photon events; geometry replicated); per-channel DAQ
accumulators are combined with one NCCL reduction.

--impl reference runs the REFERENCE's own CUDA kernels (oracle/_ref/*.cubin,
compiled from /root/reference by oracle/Makefile) through a launch-for-launch
replay of chroma/gpu/photon.py:240-290 on the same inputs: the reference has no
CPU propagator, so this is "the reference's implementation of the path"
(BASELINE.md section 2).  Rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

MAX_STEPS = 100
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel from the
# committed `ncu --set full` capture (profiles/), keyed by (workload, photons per event)
NCU_TRAFFIC_BYTES_PER_LAUNCH = {
    # profiles/r01_ncu_step_intersect_summary.txt: 2.678675 GB read + 47.112192 MB written
    ('pmt29k', 2500000): 2678675000 + 47112192,
}


_REAL_STDOUT = None


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def emit(obj):
    data = (json.dumps(obj) + '\n').encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


# ------------------------------------------------------------------ clocks
class ClockSampler(object):
    """SM clock and throttle reasons sampled DURING the timed regions.  NVML in a thread
    (2 ms period; the timed region of the default run is only ~40 ms, shorter than an
    nvidia-smi process takes to start), nvidia-smi -lms as the fallback."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self.nvml = index, [], None, None
        self.sm, self.reasons, self.sm_max = [], set(), None
        self._stop = threading.Event()
        self._on = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get('CUDA_VISIBLE_DEVICES', '')
            ids = [int(x) for x in vis.split(',')] if vis and all(x.strip().isdigit() for x in vis.split(',')) else None
            phys = ids[index] if ids and index < len(ids) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def start(self):
        """Start the sampling thread (idle until resume())."""
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def resume(self):
        self._on.set()

    def pause(self):
        self._on.clear()

    def _poll(self):
        nv = self.nvml
        names = ((nv.nvmlClocksThrottleReasonHwSlowdown, 'hw_slowdown'),
                 (nv.nvmlClocksThrottleReasonHwThermalSlowdown, 'hw_thermal_slowdown'),
                 (nv.nvmlClocksThrottleReasonSwThermalSlowdown, 'sw_thermal_slowdown'),
                 (nv.nvmlClocksThrottleReasonSwPowerCap, 'sw_power_cap'))
        while not self._stop.is_set():
            if not self._on.wait(0.05):
                continue
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for bit, name in names:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            if self._on.is_set():
                self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        self._stop.set()
        if self.nvml is not None:
            self.thread.join(timeout=1.0)
            return {'sm_mhz': float(np.median(self.sm)) if self.sm else None, 'sm_max_mhz': self.sm_max,
                    'reasons': sorted(self.reasons), 'samples': len(self.sm), 'source': 'nvml, 2 ms period, timed regions only'}
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 7]
        sm = [float(r[0]) for r in rows if r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace('.', '').isdigit()]
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        reasons = sorted({n for r in rows for n, v in zip(names, r[3:7]) if v.lower().startswith('active')})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(rows), 'source': 'nvidia-smi -lms 100'}


# ------------------------------------------------------------------ workload
def cache_dir():
    d = os.environ.get('CHROMA_B200_CACHE', '/tmp/chroma_b200_cache')
    os.makedirs(d, exist_ok=True)
    return d


class FlatGeometry(object):
    """A flattened detector restored from the cache (duck-types Geometry/Detector)."""
    pass


def build_detector(workload, timings):
    """Build (or load) the flattened detector + reference-format BVH."""
    from chroma_lite_b200 import demo
    from chroma_lite_b200.bvh import BVH, WorldCoords, make_recursive_grid_bvh, uint4
    from chroma_lite_b200.geometry import Mesh
    path = os.path.join(cache_dir(), 'det_%s_v2.npz' % workload)
    t0 = time.perf_counter()
    if workload == 'scint':
        # BASELINE config 4: re-emitting scintillator in an acrylic vessel, WLS shell, dichroic /
        # angular / thin-film surfaces (tests/scenes.py; no reference fixture exists).
        # Small mesh, built in a second: no cache.
        import scenes
        det = scenes.scintillator_scene(96)
        timings.update(author_s=time.perf_counter() - t0, flatten_s=0.0, bvh_s=0.0, cached=False)
        return det
    if workload == 'pmt29k':
        det = demo.detector_29k()
    elif workload == 'tiny':
        det = demo.tiny()
    else:
        raise SystemExit('unknown workload ' + workload)
    timings['author_s'] = time.perf_counter() - t0
    if os.path.exists(path):
        z = np.load(path)
        det.mesh = Mesh.__new__(Mesh)
        det.mesh.vertices, det.mesh.triangles = z['vertices'], z['triangles']
        det.colors, det.solid_id = z['colors'], z['solid_id']
        det.material1_index, det.material2_index, det.surface_index = z['m1'], z['m2'], z['surf']
        # material/surface object lists in the same order the cache was written with
        det.flatten_objects_only = True
        _restore_object_lists(det)
        det.solid_id_to_channel_index = np.asarray(det.solid_id_to_channel_index, dtype=np.int32)
        det.bvh = BVH(WorldCoords(z['world_origin'], z['world_scale']), z['nodes'].view(uint4)[:, 0], z['layers'])
        timings['flatten_s'] = float(z['flatten_s'])
        timings['bvh_s'] = float(z['bvh_s'])
        timings['cached'] = True
        return det
    t0 = time.perf_counter()
    det.flatten()                       # incl. global vertex de-duplication, as the reference does
    timings['flatten_s'] = time.perf_counter() - t0
    t0 = time.perf_counter()
    det.bvh = make_recursive_grid_bvh(det.mesh)
    timings['bvh_s'] = time.perf_counter() - t0
    timings['cached'] = False
    try:
        tmp = path + '.tmp%d.npz' % os.getpid()
        np.savez(tmp, vertices=det.mesh.vertices, triangles=det.mesh.triangles, colors=det.colors,
                 solid_id=det.solid_id, m1=det.material1_index, m2=det.material2_index, surf=det.surface_index,
                 world_origin=det.bvh.world_coords.world_origin, world_scale=det.bvh.world_coords.world_scale,
                 nodes=det.bvh.nodes.view(np.uint32).reshape(-1, 4), layers=np.asarray(det.bvh.layer_offsets),
                 flatten_s=timings['flatten_s'], bvh_s=timings['bvh_s'])
        os.replace(tmp, path)
    except Exception as e:               # cache is best effort
        log('cache write failed:', e)
    return det


def _restore_object_lists(det):
    from chroma_lite_b200.geometry import _unique_objects
    det.unique_materials = _unique_objects([m for s in det.solids for m in s.unique_materials])
    det.unique_surfaces = _unique_objects([x for s in det.solids for x in s.unique_surfaces])


METRIC_NAME = {'scint': 'photons propagated/sec on the liquid-scintillator detector (config 4)'}
WL_RANGE = {'scint': (250.0, 450.0)}      # per workload; default 300-600 nm
_workload = 'pmt29k'


def make_event(n, seed):
    import scenes
    return scenes.point_source(n, seed=seed, wl_range=WL_RANGE.get(_workload, (300.0, 600.0)))


def algorithmic_bytes(det, desc, sample_photons, timings):
    """B_photon = 120 + S*(16*Nnode + 48*Ntri + 64) with Nnode/Ntri/S measured by the
    oracle's reference-order traversal on a bounded sample; also times
    the oracle = the cpu_baseline ("port", all host cores: photons are independent)."""
    from oracle import orc
    t0 = time.perf_counter()
    st = orc.rng_init(42, 0, len(sample_photons))
    t1 = time.perf_counter()
    bank, cnt = orc.propagate(desc, sample_photons, st, max_steps=MAX_STEPS)
    t2 = time.perf_counter()
    n = len(sample_photons)
    nnode = cnt['nodes'] / max(cnt['calls'], 1)
    ntri = cnt['tris'] / max(cnt['calls'], 1)
    steps = cnt['steps'] / n
    b = 120.0 + steps * (16.0 * nnode + 48.0 * ntri + 64.0)
    timings['oracle'] = {'photons': n, 'seconds': t2 - t1, 'rng_init_s': t1 - t0, 'threads': orc.threads(),
                         'nodes_per_call': nnode,
                         'tris_per_call': ntri, 'steps_per_photon': steps, 'bytes_per_photon': b,
                         'max_stack': cnt['max_stack']}
    return b, n / (t2 - t1)


def dist_setup(ngpus):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()


def max_over_ranks(x, world):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, world):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


# ------------------------------------------------------------------ our arm
def run_ours(args):
    import ctypes as C
    from chroma_lite_b200 import gpu, sim, _lib, parallel, event
    from chroma_lite_b200.gpu.geometry import make_desc
    rank, world, local = dist_setup(args.gpus)
    _lib.init(local)
    lib = _lib.lib()
    timings = {}
    det = build_detector(args.workload, timings)
    t0 = time.perf_counter()
    s = sim.Simulation(det, seed=42 + rank, cuda_device=local, nthreads_per_block=512,
                       max_blocks=max(1024, -(-args.photons // 512)))
    timings['upload_geometry_s'] = time.perf_counter() - t0
    g, rng = s.gpu_geometry, s.rng_states
    n = args.photons
    ev = make_event(n, seed=1000 + rank)
    gp = gpu.GPUPhotons(ev)
    pristine = gpu.GPUPhotons(ev)
    fields = ('pos', 'dir', 'pol', 'wavelengths', 't', 'last_hit_triangles', 'flags', 'weights', 'evidx')

    def restore():
        for f in fields:
            getattr(gp, f).copy_from_device(getattr(pristine, f).ptr)

    def one_step():
        restore()
        lib.cb_flush_l2()
        gp.propagate(g, rng, nthreads_per_block=512, max_blocks=s.max_blocks, max_steps=MAX_STEPS)
        return gp.last_stats

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        one_step()
    barrier(world)
    _lib.check(lib.cb_synchronize())
    sampler.resume()
    kernel_ms, launches, steps_taken = 0.0, 0, 0
    nodes_v, tris_v, resolved = 0, 0, 0
    int0_ms, int0_rays, int0_n = 0.0, 0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st = one_step()
        kernel_ms += st.kernel_ms
        launches += st.launches
        steps_taken += st.steps
        nodes_v += st.nodes_visited
        tris_v += st.tris_tested
        resolved += st.rays_resolved
        if st.intersect0_rays:
            int0_ms += st.intersect0_ms
            int0_rays += st.intersect0_rays
            int0_n += 1
    _lib.check(lib.cb_synchronize())
    barrier(world)
    wall = time.perf_counter() - t0
    sampler.pause()
    # device time of the propagate kernels (CUDA events on the launching stream), max over ranks
    dev_s = max_over_ranks(kernel_ms / 1e3, world)
    total_photons = sum_over_ranks(float(n * args.steps), world)
    value = total_photons / dev_s

    # ---- e2e through the public API: K events as HOST arrays in, flat hits + DAQ channels
    # out (Simulation.simulate, batches double-buffered), then ONE reduction of the
    # run-level per-channel hit counts / charge across ranks
    import torch
    h2d = sum(getattr(ev, f).nbytes for f in ('pos', 'dir', 'pol', 'wavelengths', 't', 'flags', 'evidx'))
    sim_kw = dict(keep_hits=False, keep_flat_hits=True, run_daq=True, max_steps=MAX_STEPS, photons_per_batch=n)
    # the event's host arrays live in page-locked memory (gpu.pagelocked_empty, the role of
    # pycuda's pagelocked_empty in the reference): every step uploads them again, host -> device
    ev = gpu.pin_photons(ev)
    list(s.simulate((event.Event(photons_beg=ev) for _ in range(max(4, args.warmup))), **sim_kw))   # warm-up: 3 batches in flight
    nch = s.gpu_geometry.nchannels
    barrier(world)
    _lib.check(lib.cb_synchronize())
    sampler.resume()
    t0 = time.perf_counter()
    hit_count = np.zeros(nch, dtype=np.int64)
    charge = np.zeros(nch, dtype=np.float64)
    d2h = 0
    for out_ev in s.simulate((event.Event(photons_beg=ev) for _ in range(args.steps)), **sim_kw):
        hit_count += out_ev.channels.hit
        charge += np.where(out_ev.channels.hit, out_ev.channels.q, 0.0)
        fh = out_ev.flat_hits
        d2h = sum(getattr(fh, f).nbytes for f in fields) + fh.channel.nbytes + 3 * 4 * nch
    if world > 1:
        import torch.distributed as dist
        buf = torch.from_numpy(np.concatenate([hit_count.astype(np.float64), charge])).cuda()
        dist.reduce(buf, dst=0, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()
    _lib.check(lib.cb_synchronize())
    barrier(world)
    e2e_s = max_over_ranks(time.perf_counter() - t0, world)
    clocks = sampler.stop() if rank == 0 else None
    e2e = total_photons / e2e_s
    timings['e2e_last_batch'] = dict(s.last_timings)
    timings['e2e_s_per_event'] = e2e_s / args.steps
    timings['e2e_hits_per_event'] = int(len(fh))

    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    # ---- roofline + cpu baseline (rank 0, N=1 only does the CPU leg)
    roofline, cpu_baseline = None, None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    if world == 1:
        from oracle import orc
        desc, keep = make_desc(det)
        b_photon, cpu_rate = algorithmic_bytes(det, desc, make_event(args.cpu_sample, seed=999), timings)
        # dominant kernel = the first step's traversal kernel (one ray per photon).  Algorithmic
        # bytes per ray B_ray = 32 + 16*Nnode + 48*Ntri with Nnode/Ntri counted by the REFERENCE
        # traversal (oracle, reference tree) on a sample of the same rays
        smp = make_event(min(args.cpu_sample, 20000), seed=998)
        _, _, c0 = orc.intersect(desc, smp.pos, smp.dir)
        b_ray = 32.0 + 16.0 * c0['nodes'] / len(smp) + 48.0 * c0['tris'] / len(smp)
        timings['oracle']['first_step_nodes_per_ray'] = c0['nodes'] / len(smp)
        timings['oracle']['first_step_tris_per_ray'] = c0['tris'] / len(smp)
        if int0_n:
            per_launch_s = int0_ms / 1e3 / int0_n
            rays_per_launch = int0_rays / int0_n
            achieved = b_ray * rays_per_launch / per_launch_s / 1e9
            roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                        'traffic': NCU_TRAFFIC_BYTES_PER_LAUNCH.get((args.workload, n)),
                        'kernel': 'step_intersect_kernel (first step)', 'bytes_per_ray': b_ray,
                        'rays_per_launch': rays_per_launch, 'ms_per_launch': per_launch_s * 1e3,
                        'rays_per_s': rays_per_launch / per_launch_s, 'bytes_per_photon_all_steps': b_photon,
                        'peak_source': 'MEASURED_PEAKS.json hbm_gbs (burst copy)' if peaks else 'fallback 6650 GB/s'}
        cpu_baseline = {'value': cpu_rate, 'unit': 'photons/s', 'cores': orc.threads(), 'kind': 'port',
                        'sample': '%d photons of the same event type through oracle/chroma_oracle.c '
                                  '(orc_propagate, max_steps=%d) on %d host threads, host cores on this box: %d'
                                  % (args.cpu_sample, MAX_STEPS, orc.threads(), os.cpu_count())}
    line = {
        'metric': METRIC_NAME.get(args.workload, 'photons propagated/sec (whole box) on 29k-PMT detector'), 'value': value,
        'unit': 'photons/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dev_s * 1e3 / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': args.workload, 'photons_per_event': n, 'max_steps': MAX_STEPS,
                   'triangles': int(len(det.mesh.triangles)), 'bvh_nodes': int(len(det.bvh.nodes)),
                   'channels': int(det.num_channels()), 'rng_pool': int(len(rng)),
                   'l2': 'flushed between steps (cb_flush_l2 writes 2x L2) and node+triangle arrays exceed L2',
                   'parallelism': 'photon banks sharded x%d, geometry replicated' % world,
                   'tree': '%s, leaf split %s' % (os.environ.get('CHROMA_B200_TREE') or 'solids first',
                                                  os.environ.get('CHROMA_B200_LEAF_SPLIT') or 'off')},
        'e2e': {'value': e2e, 'unit': 'photons/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h)},
        'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu_baseline,
        'extra': {'steps_per_photon': steps_taken / float(n * args.steps), 'wall_s': wall, 'setup': timings,
                  'engine_counters': {'entries_per_traversal': nodes_v / max(steps_taken, 1),
                                      'tris_per_traversal': tris_v / max(steps_taken, 1),
                                      'rays_redone_fraction': resolved / max(steps_taken, 1),
                                      'enabled': bool(os.environ.get('CHROMA_B200_STATS'))},
                  'host_cores': os.cpu_count()},
    }
    emit(line)


# ------------------------------------------------------------------ ray microbench (BASELINE config 2)
def rays_scene():
    """~1.2 M triangles: a finely tessellated sphere shell around a coarser one (the ~1M-triangle
    STL BASELINE names is not shipped with the reference)."""
    from chroma_lite_b200.geometry import Geometry, Solid, vacuum
    from chroma_lite_b200.make import sphere
    from chroma_lite_b200.demo import optics
    geo = Geometry(optics.water)
    geo.add_solid(Solid(sphere(1000.0, 708), optics.glass, optics.water))
    geo.add_solid(Solid(sphere(600.0, 300), vacuum, optics.glass))
    geo.flatten()
    return geo


def make_rays(geo, n, seed=1234):
    rng = np.random.default_rng(seed)
    lo, hi = geo.mesh.vertices.min(axis=0), geo.mesh.vertices.max(axis=0)
    c, h = (lo + hi) / 2, (hi - lo) / 2 * 1.5
    o = (c + h * rng.uniform(-1, 1, (n, 3))).astype(np.float32)
    u = rng.uniform(-1, 1, n)
    phi = rng.uniform(0, 2 * np.pi, n)
    s = np.sqrt(1 - u * u)
    d = np.column_stack([s * np.cos(phi), s * np.sin(phi), u]).astype(np.float32)
    return o, d


def run_rays(args):
    """rays/s of the nearest-hit query (triangle index + distance) on a ~1.2 M-triangle mesh."""
    from chroma_lite_b200 import gpu, _lib
    from chroma_lite_b200 import gpuarray as ga
    from chroma_lite_b200.gpu.tools import to_float3
    from chroma_lite_b200.bvh import make_recursive_grid_bvh
    _lib.init(0)
    lib = _lib.lib()
    n = args.photons if args.photons != 2500000 else 10000000
    geo = rays_scene()
    t0 = time.perf_counter()
    geo.bvh = make_recursive_grid_bvh(geo.mesh)
    bvh_s = time.perf_counter() - t0
    o, d = make_rays(geo, n)
    line = {'metric': 'rays/s, nearest-hit triangle + distance through the BVH', 'unit': 'rays/s', 'n_gpus': 1,
            'steps': args.steps, 'warmup': args.warmup, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'rays', 'rays': n, 'triangles': int(len(geo.mesh.triangles)),
                       'bvh_nodes': int(len(geo.bvh.nodes)), 'bvh_build_s': bvh_s,
                       'l2': 'ray arrays (320 MB) and geometry exceed L2'}}
    if args.impl == 'reference':
        from oracle import ref_driver
        from chroma_lite_b200.gpu.geometry import make_desc
        desc, keep = make_desc(geo)
        rg = ref_driver.RefGeometry(desc, keep)
        ms = [ref_driver.intersect(rg, o, d, block=64)[2] for _ in range(args.warmup + args.steps)][args.warmup:]
        line.update(impl='reference', value=n * len(ms) / (sum(ms) / 1e3), ms_per_step=sum(ms) / len(ms),
                    e2e={'value': n * len(ms) / (sum(ms) / 1e3), 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                    cpu_baseline={'value': n * len(ms) / (sum(ms) / 1e3), 'unit': 'rays/s', 'cores': 1, 'kind': 'reference',
                                  'sample': 'reference intersect_mesh (oracle/_ref/ref_wrap.cubin) on one B200, 64-thread blocks'})
        emit(line)
        return
    g = gpu.GPUGeometry(geo)
    do, dd = ga.to_gpu(to_float3(o)), ga.to_gpu(to_float3(d))
    # the scene build above leaves the GPU idle for seconds: warm up until the clocks are back up
    # (the first calls after an idle period run 2x slower), at least 30 calls / 0.15 s
    t_w, n_w = time.perf_counter(), 0
    while n_w < max(args.warmup, 30) or time.perf_counter() - t_w < 0.15:
        gpu.intersect_mesh(g, do, dd)
        n_w += 1
    ms = []
    for _ in range(args.steps):
        lib.cb_flush_l2()
        _lib.check(lib.cb_synchronize())
        _lib.check(lib.cb_timer_start())
        tri, dist = gpu.intersect_mesh(g, do, dd)
        t = _lib.C.c_float()
        _lib.check(lib.cb_timer_stop(_lib.C.byref(t)))
        ms.append(t.value)
    # e2e: host arrays in (page-locked), triangle + distance back on the host
    po, pd = gpu.pagelocked_copy(o), gpu.pagelocked_copy(d)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tri, dist = gpu.intersect_mesh(g, po, pd)
        ht, hd = tri.get(), dist.get()
    e2e_s = time.perf_counter() - t0
    line.update(value=n * len(ms) / (sum(ms) / 1e3), ms_per_step=sum(ms) / len(ms), gpu_launches=args.steps,
                e2e={'value': n * args.steps / e2e_s, 'unit': 'rays/s', 'h2d_bytes_per_step': int(o.nbytes + d.nbytes),
                     'd2h_bytes_per_step': int(ht.nbytes + hd.nbytes)},
                extra={'hit_fraction': float((ht >= 0).mean())})
    emit(line)


# ------------------------------------------------------------------ PDF accumulators
def run_pdf(args):
    """channel-copies/s of GPUPDF.accumulate_pdf_eval on the 29k-PMT detector's channel count:
    one step = one acquisition of ndaq DAQ copies (device-resident times) merged into the
    per-channel nearest-neighbour lists.  --impl reference: the reference's accumulate_bincount +
    accumulate_nearest_neighbor_block with its work-queue fill and two synchronisations."""
    from chroma_lite_b200 import gpu, _lib
    from chroma_lite_b200 import gpuarray as ga
    _lib.init(0)
    nch, ndaq, m = 28995, 64, 100
    rng = np.random.default_rng(7)
    event_hit = rng.uniform(size=nch) < 0.35
    event_time = rng.normal(60.0, 8.0, nch).astype(np.float32)
    acqs = []
    for k in range(4):
        t = rng.normal(60.0, 10.0, (ndaq, nch)).astype(np.float32)
        t[rng.uniform(size=(ndaq, nch)) > 0.3] = 1e9
        acqs.append(t.reshape(-1))
    line = {'metric': 'channel-copies/s accumulated into per-channel PDF evaluations', 'unit': 'channel-copies/s', 'n_gpus': 1,
            'steps': args.steps, 'warmup': args.warmup, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'pdf', 'channels': nch, 'ndaq': ndaq, 'min_bin_content': m, 'hit_channels': int(event_hit.sum()),
                       'l2': 'working set (7.4 MB of times per acquisition) is L2-resident by nature of the workload'}}
    if args.impl == 'reference':
        from oracle import ref_driver
        ref = ref_driver.RefPDF()
        ref.setup_pdf_eval(event_hit, event_time, 2.0, (0.0, 200.0), min_bin_content=m)
        dev = [ref_driver.to_dev(a) for a in acqs]
        t_w, k = time.perf_counter(), 0
        while k < args.warmup or time.perf_counter() - t_w < 0.3:      # until the clocks are back up
            ref.accumulate_pdf_eval(dev[k % 4], ndaq)
            k += 1
        ref.setup_pdf_eval(event_hit, event_time, 2.0, (0.0, 200.0), min_bin_content=m)    # same state as our arm starts from
        timer = ref_driver.Timer()
        dev_ms, t0 = 0.0, time.perf_counter()
        for k in range(args.steps):
            dev_ms += ref.accumulate_pdf_eval(dev[k % 4], ndaq, timer=timer)
        dt = time.perf_counter() - t0
        v = nch * ndaq * args.steps / (dev_ms / 1e3)
        line.update(impl='reference', value=v, ms_per_step=dev_ms / args.steps,
                    e2e={'value': nch * ndaq * args.steps / dt, 'unit': 'channel-copies/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                    cpu_baseline={'value': v, 'unit': 'channel-copies/s', 'cores': 1, 'kind': 'reference',
                                  'sample': 'reference pdf.cu kernels (oracle/_ref/pdf.cubin) on one B200: work-queue fill + '
                                            'accumulate_bincount + accumulate_nearest_neighbor_block, CUDA-event time'})
        emit(line)
        return
    lib = _lib.lib()
    p = gpu.GPUPDF()
    p.setup_pdf_eval(event_hit, event_time, event_time, 2.0, (0.0, 200.0), 1.0, (0.0, 10.0), min_bin_content=m)
    chans = [gpu.GPUChannels(ga.to_gpu(a), ga.to_gpu(a), ga.to_gpu(np.zeros(len(a), np.uint32)), ndaq=ndaq, stride=nch) for a in acqs]
    t_w, k = time.perf_counter(), 0
    while k < args.warmup or time.perf_counter() - t_w < 0.3:          # until the clocks are back up
        p.accumulate_pdf_eval(chans[k % 4])
        k += 1
    p.clear_pdf_eval()                                                   # both arms time the first K acquisitions of a fresh evaluation
    _lib.check(lib.cb_synchronize())
    _lib.check(lib.cb_timer_start())
    t0 = time.perf_counter()
    for k in range(args.steps):
        p.accumulate_pdf_eval(chans[k % 4])
    tms = _lib.C.c_float()
    _lib.check(lib.cb_timer_stop(_lib.C.byref(tms)))
    wall = time.perf_counter() - t0
    line.update(value=nch * ndaq * args.steps / (tms.value / 1e3), ms_per_step=tms.value / args.steps, gpu_launches=args.steps,
                e2e={'value': nch * ndaq * args.steps / wall, 'unit': 'channel-copies/s', 'h2d_bytes_per_step': 0,
                     'd2h_bytes_per_step': 0, 'note': 'host-timed through GPUPDF.accumulate_pdf_eval; the DAQ output is device-resident by construction'})
    emit(line)


# ------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import ref_driver
    if not ref_driver.available():
        emit({'impl': 'reference', 'unavailable': 'oracle/_ref cubins missing (build() needs /root/reference)'})
        return
    from chroma_lite_b200 import _lib
    from chroma_lite_b200.gpu.geometry import make_desc
    _lib.init(0)                         # only for the BVH builder / cache; no engine kernel is timed below
    timings = {}
    det = build_detector(args.workload, timings)
    desc, keep = make_desc(det)
    t0 = time.perf_counter()
    rg = ref_driver.RefGeometry(desc, keep)
    timings['upload_geometry_s'] = time.perf_counter() - t0
    n = args.photons
    ev = make_event(n, seed=1000)
    rng = ref_driver.RefRNG(512 * 1024, seed=42)        # Simulation defaults (chroma/sim.py:23-24,50)
    sampler = ClockSampler(0)

    rg.attach_detector(det)

    def one_step():
        # chroma/sim.py:54-154 for one event: upload, propagate, flat hits, DAQ, channels
        rp = ref_driver.RefPhotons(ev)
        r = rp.propagate(rg, rng, nthreads_per_block=512, max_blocks=1024, max_steps=MAX_STEPS)
        hits = rp.get_flat_hits(rg)
        ch = ref_driver.run_daq(rg, rp, rng, nthreads_per_block=512, max_blocks=1024)
        return rp, r, hits, ch

    sampler.start()
    for _ in range(args.warmup):
        one_step()
    sampler.resume()
    ms, launches = 0.0, 0
    t_e2e = 0.0
    for _ in range(args.steps):
        ref_driver.sync()
        t0 = time.perf_counter()
        rp, r, hits, ch = one_step()
        ref_driver.sync()
        t_e2e += time.perf_counter() - t0
        ms += r['ms']
        launches += r['launches']
    out = rp.get()
    clocks = sampler.stop()
    value = n * args.steps / (ms / 1e3)
    e2e = n * args.steps / t_e2e
    line = {
        'impl': 'reference', 'metric': METRIC_NAME.get(args.workload, 'photons propagated/sec (whole box) on 29k-PMT detector'),
        'value': value,
        'unit': 'photons/s', 'n_gpus': 1, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': args.workload, 'photons_per_event': n, 'max_steps': MAX_STEPS,
                   'triangles': int(len(det.mesh.triangles)), 'bvh_nodes': int(len(det.bvh.nodes)),
                   'launch': '512 threads x 1024 blocks, rng pool 524288, host queue loop (chroma/gpu/photon.py:240-290)'},
        'cpu_baseline': {'value': value, 'unit': 'photons/s', 'cores': 1, 'kind': 'reference',
                         'sample': 'reference CUDA kernels (oracle/_ref/propagate.cubin, sm_100a, reference nvcc flags) '
                                   'on one B200 driven by one host thread; the reference has no CPU propagator'},
        'e2e': {'value': e2e, 'unit': 'photons/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': int(launches), 'clocks': clocks,
        'extra': {'setup': timings, 'host_cores': os.cpu_count(),
                  'terminal_fraction': float(((out.flags & 0x800F) != 0).mean())},
    }
    emit(line)


def main():
    # stdout carries exactly ONE JSON line: anything libraries print there (NCCL's
    # version banner, ...) is diverted to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='pmt29k')
    ap.add_argument('--photons', type=int, default=2500000)
    ap.add_argument('--cpu-sample', type=int, default=1000000)
    args = ap.parse_args()
    global _workload
    _workload = args.workload
    if args.workload == 'scint' and args.photons == 2500000:
        args.photons = 10000000            # config 4 is quoted at 10 M photons per event
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.workload == 'rays':
        run_rays(args)
    elif args.workload == 'pdf':
        run_pdf(args)
    elif args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
