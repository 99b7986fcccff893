"""XORWOW parity on the GPU: engine vs oracle (bit-exact) and vs cuRAND's own
curand_init/curand compiled from the toolkit header (oracle/_ref)."""
import numpy as np
import pytest

from chroma_lite_b200 import gpu
from chroma_lite_b200 import gpuarray as ga
from chroma_lite_b200 import _lib
from oracle import orc, ref_driver

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('seed', [0, 1, 2 ** 32 + 5])
def test_states_match_oracle(gpu_ready, seed):
    n = 5000
    rng = gpu.get_rng_states(n, seed=seed)
    got = rng.get()
    idx = np.r_[0:64, 1000:1010, n - 3:n]
    for i in idx:
        assert np.array_equal(got[i], orc.rng_init(seed, int(i), 1)[0]), 'stream %d' % i


def test_states_match_curand_init(gpu_ready):
    n = 3000
    for seed, first in ((1, 0), (2 ** 32 + 5, 0)):
        words, st = ref_driver.rng_words(n, seed, first_stream=first, ndraw=3)
        rng = gpu.get_rng_states(n, seed=seed)
        assert np.array_equal(rng.get(), st)
    # high stream ids (cuRAND applies more skip matrices): 524287 and 10^8
    for first in (524287, 10 ** 8):
        words, st = ref_driver.rng_words(8, 1, first_stream=first, ndraw=3)
        mine = orc.rng_init(1, first, 8)
        assert np.array_equal(mine, st)
        assert np.array_equal(orc.rng_words(mine.copy(), 3), words)


def test_fill_uniform_matches_oracle_and_advances(gpu_ready):
    n = 4096
    rng = gpu.get_rng_states(n, seed=9)
    out = ga.empty(n, np.float32)
    for rep in range(2):
        _lib.check(_lib.lib().cb_rng_fill_uniform(rng.handle, n, 0.0, 1.0, out.ptr))
    st = orc.rng_init(9, 0, n)
    orc.rng_fill_uniform(st, 0.0, 1.0)
    want = orc.rng_fill_uniform(st, 0.0, 1.0)
    assert np.array_equal(out.get(), want)          # curand_uniform itself: bit-exact
    assert np.array_equal(rng.get(), st)
    _lib.check(_lib.lib().cb_rng_fill_uniform(rng.handle, n, -1.0, 2.0, out.ptr))
    want = orc.rng_fill_uniform(st, -1.0, 2.0)      # low + u*(high-low) is FMA-contracted on the device
    assert np.allclose(out.get(), want, rtol=0, atol=3e-7) and np.array_equal(rng.get(), st)
