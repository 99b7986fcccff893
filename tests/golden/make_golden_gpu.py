"""Generate tests/golden/ref_kernel_histories.npz: end states of the REFERENCE's own propagate and
run_daq kernels (oracle/_ref/*.cubin: propagate.cu / daq.cu compiled where they lie with the
reference's nvcc flags, driven by oracle/ref_driver.py) on the small cases of ref_kernel_cases.py.

Needs a GPU, not /root/reference:   gpurun -- 'python tests/golden/make_golden_gpu.py gpurun_out'
then copy gpurun_out/ref_kernel_histories.npz to tests/golden/.  The fixture pins the physics of the
CPU oracle (oracle/chroma_oracle.c) in the GPU-less test tier (tests/test_oracle_physics.py)."""
import os
import sys
import traceback
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE), HERE]

import scenes                              # noqa: E402
from oracle import ref_driver              # noqa: E402
from ref_kernel_cases import CASES, build  # noqa: E402


def main(outdir):
    out = {}
    for name, c in CASES.items():
        try:
            geo, ph = build(name)
            desc, keep = scenes.desc_of(geo)
            rg = ref_driver.RefGeometry(desc, keep)
            rng = ref_driver.RefRNG(c['n'], seed=c['rng_seed'])
            rp = ref_driver.RefPhotons(ph)
            rp.propagate(rg, rng, nthreads_per_block=256, max_steps=c['max_steps'], use_weights=c['use_weights'],
                         scatter_first=c['scatter_first'], force_single_launch=True)
            end = rp.get()
            out[name + '.input_dir_sum'] = np.asarray(ph.dir, dtype=np.float32).sum(axis=0, dtype=np.float64)
            out[name + '.ntriangles'] = np.int64(len(geo.mesh.triangles))
            for f in ('pos', 'dir', 'pol', 'wavelengths', 't', 'weights'):
                out['%s.%s' % (name, f)] = np.asarray(getattr(end, f), dtype=np.float32)
            out[name + '.flags'] = np.asarray(end.flags, dtype=np.uint32)
            out[name + '.last_hit_triangles'] = np.asarray(end.last_hit_triangles, dtype=np.int32)
            out[name + '.rng'] = rng.states6()
            if 'daq_seed' in c:
                rg.attach_detector(geo)
                drng = ref_driver.RefRNG(c['n'], seed=c['daq_seed'])
                t, q, hist, tint, qint = ref_driver.run_daq(rg, rp, drng, nthreads_per_block=64,
                                                            max_blocks=(c['n'] + 63) // 64)
                out[name + '.daq_t'], out[name + '.daq_q'], out[name + '.daq_flags'] = t, q, hist
                out[name + '.daq_time_int'], out[name + '.daq_q_int'] = tint, qint
            print(name, 'ok', flush=True)
        except Exception:
            print(name, 'FAILED', flush=True)
            traceback.print_exc()
    os.makedirs(outdir, exist_ok=True)
    np.savez_compressed(os.path.join(outdir, 'ref_kernel_histories.npz'), **out)
    print('wrote', len(out), 'arrays')


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else HERE)
