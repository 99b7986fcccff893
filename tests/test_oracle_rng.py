"""Pins the oracle's XORWOW against cuRAND itself (no GPU needed):
libcurand's HOST generator and the toolkit's precalculated skip matrices."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest

from oracle import orc

CURAND = '/usr/local/cuda/lib64/libcurand.so'
PRECALC = '/usr/local/cuda/include/curand_precalc.h'


@pytest.mark.skipif(not os.path.exists(CURAND), reason='libcurand not installed')
@pytest.mark.parametrize('seed', [0, 1, 2 ** 32 + 5, 0xDEADBEEFCAFE])
def test_oracle_matches_curand_host_generator(seed):
    # XORWOW host generator, LEGACY ordering: output i comes from stream i % 4096
    cur = C.CDLL(CURAND)
    gen = C.c_void_p()
    assert cur.curandCreateGeneratorHost(C.byref(gen), 101) == 0
    assert cur.curandSetGeneratorOrdering(gen, 103) == 0
    assert cur.curandSetPseudoRandomGeneratorSeed(gen, C.c_ulonglong(seed)) == 0
    ndraw = 5
    out = np.zeros(4096 * ndraw, dtype=np.uint32)
    assert cur.curandGenerate(gen, out.ctypes.data_as(C.c_void_p), C.c_size_t(out.size)) == 0
    cur.curandDestroyGenerator(gen)
    ref = out.reshape(ndraw, 4096).T           # [stream, draw]
    streams = np.array([0, 1, 2, 3, 63, 64, 1000, 4095])
    for s in streams:
        st = orc.rng_init(seed, int(s), 1)
        assert np.array_equal(orc.rng_words(st, ndraw)[0], ref[s]), 'stream %d' % s


@pytest.mark.skipif(not os.path.exists(PRECALC), reason='curand_precalc.h not installed')
def test_skip_matrices_match_curand_precalc(tmp_path):
    src = tmp_path / 'dump.c'
    src.write_text('#include <stdio.h>\n#define CURAND_XORWOW_PRECALCULATED_HOST_QUALIFIERS static\n'
                   '#define CURAND_XORWOW_PRECALCULATED_DEVICE_QUALIFIERS static\n'
                   '#define __constant__\n#define __device__\n'
                   '#include "%s"\n'
                   'int main(){fwrite(precalc_xorwow_matrix_host,4,32*800,stdout);'
                   'fwrite(precalc_xorwow_offset_matrix_host,4,32*800,stdout);return 0;}\n' % PRECALC)
    exe = tmp_path / 'dump'
    r = subprocess.run(['gcc', '-O0', '-w', '-o', str(exe), str(src)], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip('curand_precalc.h does not compile standalone: ' + r.stderr[:200])
    raw = np.frombuffer(subprocess.run([str(exe)], capture_output=True).stdout, dtype=np.uint32)
    seq, off = raw[:32 * 800].reshape(32, 800), raw[32 * 800:].reshape(32, 800)
    for k in (0, 1, 2, 7, 15, 31):
        assert np.array_equal(orc.xorwow_matrix(0, k), seq[k]), 'sequence matrix %d' % k
        assert np.array_equal(orc.xorwow_matrix(1, k), off[k]), 'offset matrix %d' % k


def test_offset_equals_stepping():
    st = orc.rng_init(7, 3, 1)
    orc.rng_words(st, 1000)
    st2 = orc.rng_init(7, 3, 1, offset=1000)
    assert np.array_equal(st, st2)


def test_uniform_range_and_fill():
    st = orc.rng_init(1, 0, 2000)
    u = orc.rng_fill_uniform(st)
    assert u.min() > 0.0 and u.max() <= 1.0
    assert abs(u.mean() - 0.5) < 0.03
    # stream 0 of seed 1: first word is fixed forever
    w = orc.rng_words(orc.rng_init(1, 0, 1), 1)[0, 0]
    assert np.float32(w) * np.float32(2.3283064e-10) + np.float32(2.3283064e-10 / 2) == u[0]
