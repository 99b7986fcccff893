"""The engine against the COMMITTED end states of the reference's kernels (tests/golden/
ref_kernel_histories.npz, made by tests/golden/make_golden_gpu.py): the same comparison as
test_gpu_propagate.py's live runs of the reference kernels, but against a fixture that does not need
oracle/_ref on the box.  Runs last (file name) so the live comparisons come first."""
import os
import sys
import numpy as np
import pytest

from chroma_lite_b200 import gpu, event
import scenes

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
from ref_kernel_cases import CASES, build   # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_kernel_histories.npz')
MIN_SAME = {'sphere': 0.998, 'tiny': 0.998, 'scint': 0.99, 'weights': 0.995, 'wires': 0.99, 'one_step': 0.999}


@pytest.mark.parametrize('name', sorted(CASES))
def test_engine_replays_reference_kernel_fixture(gpu_ready, name):
    g, c = np.load(GOLD), CASES[name]
    geo, ph = build(name)
    assert np.allclose(np.asarray(ph.dir, np.float32).sum(axis=0, dtype=np.float64), g[name + '.input_dir_sum'])
    dev = (gpu.GPUDetector if hasattr(geo, 'num_channels') else gpu.GPUGeometry)(geo)
    rng = gpu.get_rng_states(c['n'], seed=c['rng_seed'])
    gp = gpu.GPUPhotons(ph)
    gp.propagate(dev, rng, nthreads_per_block=256, max_blocks=(c['n'] + 255) // 256, max_steps=c['max_steps'],
                 use_weights=c['use_weights'], scatter_first=c['scatter_first'])
    mine = gp.get()
    flags, tri = g[name + '.flags'], g[name + '.last_hit_triangles']
    same = (mine.flags == flags) & (mine.last_hit_triangles == tri)
    assert same.mean() >= MIN_SAME[name], 'only %.5f identical histories' % same.mean()
    pos, t = g[name + '.pos'], g[name + '.t']
    scale = np.maximum(np.abs(pos[same]).max(axis=1), 1.0)
    assert (np.abs(mine.pos[same] - pos[same]).max(axis=1) / scale < 1e-4).mean() > 0.99
    assert np.isclose(mine.t[same], t[same], rtol=1e-4, atol=1e-3).mean() > 0.99
    assert (rng.get()[same] == g[name + '.rng'][same]).all(axis=1).mean() > 0.97
    if 'daq_seed' in c:
        # DAQ of the reference's end state: integer accumulators bit-exact (daq.cu:35-86)
        end = event.Photons(*[g['%s.%s' % (name, f)] for f in ('pos', 'dir', 'pol', 'wavelengths', 't',
                                                                 'last_hit_triangles', 'flags', 'weights')])
        daq = gpu.GPUDaq(dev)
        daq.begin_acquire()
        daq.acquire(gpu.GPUPhotons(end), gpu.get_rng_states(c['n'], seed=c['daq_seed']), nthreads_per_block=64,
                    max_blocks=(c['n'] + 63) // 64)
        ch = daq.end_acquire().get()
        assert np.array_equal(ch.flags, g[name + '.daq_flags'])
        assert np.array_equal(daq.channel_q_int_gpu.get(), g[name + '.daq_q_int'])
        assert np.array_equal(daq.earliest_time_int_gpu.get(), g[name + '.daq_time_int'])
        assert np.array_equal(ch.t, g[name + '.daq_t']) and np.array_equal(ch.q, g[name + '.daq_q'])
