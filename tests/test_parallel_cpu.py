"""Multi-GPU host logic on CPU: sharding and the per-channel reduction, run as a
world_size-2 gloo job (no GPU needed)."""
import os
import subprocess
import sys
import numpy as np

from chroma_lite_b200 import parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 100, 2500000):
        for w in (1, 2, 3, 8):
            r = [parallel.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    assert list(parallel.shard_events(5, 1, 2)) == [3, 4]


def test_sum_buffer_roundtrip():
    rng = np.random.default_rng(0)
    q = rng.integers(0, 2 ** 31, 100).astype(np.uint32)
    h = rng.integers(0, 2 ** 32, 100).astype(np.uint32)
    for world in (2, 8, 15, 16, 255, 256):
        q2, h2 = parallel.unpack_sum_buffer(parallel.pack_sum_buffer(q, h, world), 100, world)
        assert np.array_equal(q, q2) and np.array_equal(h, h2)


def test_or_through_sum_for_every_world_size():
    """Summing the packed per-bit counters of `world` ranks and unpacking gives the OR of their
    history words, also when every rank sets every bit (the counters must not spill over)."""
    rng = np.random.default_rng(1)
    for world in (2, 3, 8, 15, 16, 40):
        hs = [rng.integers(0, 2 ** 32, 50).astype(np.uint32) for _ in range(world)]
        hs[0][:5] = 0xFFFFFFFF
        for h in hs[1:]:
            h[:3] = 0xFFFFFFFF                      # all ranks set all bits of channels 0..2
        total = sum(parallel.pack_history(h, world).astype(np.uint64) for h in hs)
        assert total.max() < 2 ** 32               # what travels is uint32
        assert np.array_equal(parallel.unpack_history(total, world), np.bitwise_or.reduce(hs, axis=0))


def test_event_plan_streams_do_not_depend_on_the_partition():
    nev, n = 40, 2500000
    for world in (1, 2, 4, 8, 3):
        plans = [parallel.EventPlan(nev, n, r, world) for r in range(world)]
        assert sum(len(p.events) for p in plans) == nev and sum(p.nphotons for p in plans) == nev * n
        for p in plans:
            for e in p.events:
                first, count = p.window(e)
                assert count == n and p.first_stream + first == e * n      # stream of photon 0 of event e
    assert parallel.host_threads_should_block(8, 16) and not parallel.host_threads_should_block(8, 32)
    assert not parallel.host_threads_should_block(1, 16)


WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
from chroma_lite_b200 import parallel
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%%s' %% os.environ['PORT'],
                        rank=int(os.environ['RANK']), world_size=int(os.environ['WORLD_SIZE']))
rank, world = dist.get_rank(), dist.get_world_size()
C = 257
rng = np.random.default_rng(100 + rank)
t = rng.integers(1, 2**31 - 1, C); q = rng.integers(0, 2**20, C); h = rng.integers(0, 2**32, C)
rt, rq, rh = parallel.reduce_channels(torch.from_numpy(t), torch.from_numpy(q), torch.from_numpy(h))
# expected from all ranks' seeds
ts, qs, hs = [], [], []
for r in range(world):
    g = np.random.default_rng(100 + r)
    ts.append(g.integers(1, 2**31 - 1, C)); qs.append(g.integers(0, 2**20, C)); hs.append(g.integers(0, 2**32, C))
assert np.array_equal(rt.numpy(), np.min(ts, axis=0))
assert np.array_equal(rq.numpy(), np.sum(qs, axis=0))
assert np.array_equal(rh.numpy(), np.bitwise_or.reduce(hs, axis=0))
s, e = parallel.shard_range(1000, rank, world)
tot = torch.tensor([e - s]); dist.all_reduce(tot); assert tot.item() == 1000
dist.destroy_process_group()
print('rank', rank, 'ok')
'''


def test_reduce_channels_world_size_2_gloo(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % ROOT)
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE='2', PORT=port, MASTER_ADDR='127.0.0.1')
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert 'ok' in o


def test_numa_binding_reads_the_gpu_node_from_sysfs(tmp_path, monkeypatch):
    """bind_to_gpu_numa_node: cores of the GPU's node that this process may use; nothing on a host
    without NUMA information, when the node covers every usable core, or when switched off."""
    import os
    from chroma_lite_b200 import parallel
    have = sorted(os.sched_getaffinity(0))
    dev = tmp_path / '0000:66:00.0'
    dev.mkdir()
    calls = []
    monkeypatch.setattr(os, 'sched_setaffinity', lambda pid, cpus: calls.append(set(cpus)))
    (dev / 'numa_node').write_text('-1\n')
    (dev / 'local_cpulist').write_text('%d\n' % have[0])
    assert parallel.bind_to_gpu_numa_node('0000:66:00.0', sysfs=str(tmp_path)) is None and not calls
    (dev / 'numa_node').write_text('1\n')
    if len(have) > 1:
        assert parallel.bind_to_gpu_numa_node('0000:66:00.0'.upper(), sysfs=str(tmp_path)) == [have[0]]
        assert calls == [{have[0]}]
        assert parallel._cores_before_binding == len(have)
        assert parallel.host_threads_should_block(local_world_size=1) == (len(have) < 3)
    (dev / 'local_cpulist').write_text('%d-%d,100000-100003\n' % (have[0], have[-1]))
    n = len(calls)
    assert parallel.bind_to_gpu_numa_node('0000:66:00.0', sysfs=str(tmp_path)) is None and len(calls) == n   # covers everything usable
    monkeypatch.setenv('CHROMA_B200_NUMA', '0')
    (dev / 'local_cpulist').write_text('%d\n' % have[0])
    assert parallel.bind_to_gpu_numa_node('0000:66:00.0', sysfs=str(tmp_path)) is None
    assert parallel.bind_to_gpu_numa_node('0000:99:00.0', sysfs=str(tmp_path)) is None                    # no such device
    assert parallel._parse_cpulist('0-3,8,10-11') == {0, 1, 2, 3, 8, 10, 11}
