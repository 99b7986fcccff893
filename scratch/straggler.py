"""Reproduce the rank-4 straggler of the 8-GPU runs on one GPU: same seeds as rank 4 of bench.py
(Simulation seed 42 + 4, event seed 1000 + 4); prints per-event device time by kernel class."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import bench

def main():
    from chroma_lite_b200 import gpu, sim, _lib
    rank = int(os.environ.get('FAKE_RANK', '4'))
    _lib.init(0)
    lib = _lib.lib()
    os.environ.setdefault('CHROMA_B200_TREE_CACHE', bench.cache_dir())
    det = bench.build_detector('pmt29k', {})
    n = 2500000
    s = sim.Simulation(det, seed=42 + rank, cuda_device=0, nthreads_per_block=512, max_blocks=max(1024, -(-n // 512)))
    g, rng = s.gpu_geometry, s.rng_states
    ev = bench.make_event(n, seed=1000 + rank)
    gp = gpu.GPUPhotons(ev); pristine = gpu.GPUPhotons(ev)
    fields = ('pos', 'dir', 'pol', 'wavelengths', 't', 'last_hit_triangles', 'flags', 'weights', 'evidx')
    slow = None
    for k in range(int(os.environ.get('EVENTS', '12'))):
        before = rng.get()
        for f in fields:
            getattr(gp, f).copy_from_device(getattr(pristine, f).ptr)
        lib.cb_flush_l2()
        gp.propagate(g, rng, nthreads_per_block=512, max_blocks=s.max_blocks, max_steps=100)
        st = gp.last_stats
        print('event %2d kernel %.2f ms  intersect %.2f physics %.2f tail %.2f  steps %d tail photons %d tail steps %d' % (
            k, st.kernel_ms, st.intersect_ms, st.physics_ms, st.tail_ms, st.steps, st.tail_photons, st.tail_steps), flush=True)
        if st.kernel_ms > 30 and slow is None:
            slow = before
            out = gp.get()
            # which photons took many steps? flags of the end state
            d2 = (out.dir.astype(np.float64) ** 2).sum(axis=1)
            odd = np.flatnonzero(~(np.abs(d2 - 1.0) < 1e-3))
            print('photons whose direction is not unit:', len(odd))
            for i in odd[:10]:
                print('  photon', i, 'dir', out.dir[i], 'pos', out.pos[i], 'pol', out.pol[i], 'flags', hex(out.flags[i]), 'last', out.last_hit_triangles[i], 't', out.t[i])
            nanp = np.flatnonzero(np.isnan(out.pos).any(axis=1) | np.isnan(out.dir).any(axis=1) | np.isnan(out.pol).any(axis=1))
            print('photons with NaN pos/dir/pol:', len(nanp), [hex(x) for x in out.flags[nanp][:10]])
            for i in nanp[:10]:
                print('  photon', i, 'dir', out.dir[i], 'pos', out.pos[i], 'pol', out.pol[i], 'flags', hex(out.flags[i]), 'last', out.last_hit_triangles[i], 't', out.t[i])
            far = np.flatnonzero(np.abs(out.pos).max(axis=1) > 3e4)
            print('photons outside the world:', len(far), out.pos[far][:5], [hex(x) for x in out.flags[far][:5]])
    if slow is not None and os.environ.get('REPLAY'):
        # replay the slow event with per-step tracing
        rng2 = gpu.get_rng_states(len(rng), seed=1)
        import ctypes as C
        # overwrite rng2's states with the saved ones
        from chroma_lite_b200 import gpuarray as ga
        print('replaying the slow event with TRACE', flush=True)
        os.environ['CHROMA_B200_TRACE'] = '1'
        # states cannot be uploaded through the API; rerun the sequence up to that event instead
        s2 = sim.Simulation(det, seed=42 + rank, cuda_device=0, nthreads_per_block=512, max_blocks=max(1024, -(-n // 512)))
        os.environ.pop('CHROMA_B200_TRACE')
        for k in range(int(os.environ.get('EVENTS', '12'))):
            for f in fields:
                getattr(gp, f).copy_from_device(getattr(pristine, f).ptr)
            if k == int(os.environ.get('SLOW', '8')):
                os.environ['CHROMA_B200_TRACE'] = '1'
            gp.propagate(s2.gpu_geometry, s2.rng_states, nthreads_per_block=512, max_blocks=s.max_blocks, max_steps=100)
            os.environ.pop('CHROMA_B200_TRACE', None)
            if k == int(os.environ.get('SLOW', '8')):
                break

if __name__ == '__main__':
    main()
