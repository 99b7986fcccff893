"""Scratch: build the 29k-PMT scene once, then time cb_propagate under several env settings.
usage: python scratch/sweep.py "DEFER=0,TAIL=98304" "DEFER=28" ...   (TRACE=1 prints per-step times)
       CHROMA_B200_LIB variants are not switchable in-process; run once per library."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import bench


def main():
    from chroma_lite_b200 import gpu, sim, _lib
    _lib.init(0)
    lib = _lib.lib()
    workload = os.environ.get('WORKLOAD', 'pmt29k')
    n = int(os.environ.get('PHOTONS', '2500000'))
    bench._workload = workload
    det = bench.build_detector(workload, {})
    state = {'tree': None, 's': None}

    def simulation(tree):
        # the traversal tree is chosen when the geometry is created: a spec with TREE=... gets its own Simulation
        if state['s'] is None or tree != state['tree']:
            state['s'] = None
            os.environ['CHROMA_B200_TREE'] = tree
            t0 = time.perf_counter()
            state['s'] = sim.Simulation(det, seed=42, cuda_device=0, nthreads_per_block=512, max_blocks=max(1024, -(-n // 512)))
            state['tree'] = tree
            print('     geometry upload (tree %r) %.1f s' % (tree, time.perf_counter() - t0), flush=True)
        return state['s']
    s = simulation(os.environ.get('CHROMA_B200_TREE', ''))
    ev = bench.make_event(n, seed=1000)
    gp = gpu.GPUPhotons(ev); pristine = gpu.GPUPhotons(ev)
    fields = ('pos', 'dir', 'pol', 'wavelengths', 't', 'last_hit_triangles', 'flags', 'weights', 'evidx')

    def one():
        s = state['s']
        g, rng = s.gpu_geometry, s.rng_states
        for f in fields:
            getattr(gp, f).copy_from_device(getattr(pristine, f).ptr)
        lib.cb_flush_l2()
        gp.propagate(g, rng, nthreads_per_block=512, max_blocks=s.max_blocks, max_steps=100)
        return gp.last_stats

    ref_flags = None
    for spec in sys.argv[1:] or ['']:
        keys = []
        for kv in filter(None, spec.split(',')):
            k, v = kv.split('=')
            if k == 'TREE':
                simulation(v)
                continue
            os.environ['CHROMA_B200_' + k] = v
            keys.append('CHROMA_B200_' + k)
        tr = os.environ.pop('CHROMA_B200_TRACE', None)
        tl = os.environ.pop('CHROMA_B200_TIMELINE', None)
        for _ in range(3):
            one()
        ms = []
        for _ in range(5):
            lib.cb_synchronize(); t0 = time.perf_counter(); st = one(); lib.cb_synchronize(); ms.append(st.kernel_ms)
        print('     all', ' '.join('%.3f' % m for m in ms), flush=True)
        st = gp.last_stats
        print('SPEC %-40s kernel ms median %.3f min %.3f  launches %d  int0 %.3f ms' %
              (spec, float(np.median(ms)), min(ms), st.launches, st.intersect0_ms), flush=True)
        if st.nodes_visited:
            print('     stats: steps %d entries/trav %.1f tris/trav %.2f redone %.3g' % (st.steps, st.nodes_visited / st.steps,
                  st.tris_tested / st.steps, st.rays_resolved / st.steps), flush=True)
        if tr:
            os.environ['CHROMA_B200_TRACE'] = tr
            one()
            os.environ.pop('CHROMA_B200_TRACE')
        if tl:
            os.environ['CHROMA_B200_TIMELINE'] = tl
            for _ in range(3):
                st = one()
                print('     timeline run kernel_ms %.3f' % st.kernel_ms, flush=True)
            os.environ.pop('CHROMA_B200_TIMELINE')
        for k in keys:
            os.environ.pop(k, None)


if __name__ == '__main__':
    main()
