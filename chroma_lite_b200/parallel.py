"""Multi-GPU plumbing: one process per GPU, photon banks partitioned, geometry replicated,
per-channel DAQ accumulators combined with ONE reduction over NVLink.

The reference has no multi-GPU path; its atomics on the per-channel arrays
of one device (daq.cu:73-75, 143-145) are what becomes a collective once photons are sharded:
earliest_time_int -> MIN, channel_q_int -> SUM, channel_history -> bitwise OR.  NCCL has no OR,
so every bit of the history word travels as a small counter inside the SUM buffer.

The exchange itself runs inside the library (csrc/comm.cu: cb_comm_init / cb_daq_allreduce, one
grouped pair of ncclAllReduce on the library stream, float conversion fused behind it, nothing
through the host); torch.distributed is only the side channel that carries the 128-byte NCCL id
from rank 0 to the others.  `reduce_channels` is the same arithmetic on torch tensors for any
backend (the CPU tier checks it on gloo at world_size 2).

Partition invariance: give photon g (its index in the whole run) RNG stream g
(`EventPlan.first_stream`, `gpu.get_rng_states(..., first_stream=)`), and every reduced array
is bit-identical for any number of ranks: MIN, integer SUM and OR do not depend on order, and a
photon's history depends only on its own stream (tests/test_gpu_multi.py).
"""
import numpy as np


def shard_range(n, rank, world_size):
    """Contiguous [start, end) of `n` items owned by `rank` (balanced to +-1)."""
    base, rem = divmod(int(n), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_events(nevents, rank, world_size):
    """Whole events per rank so evidx and per-event DAQ stay local."""
    return range(*shard_range(nevents, rank, world_size))


class EventPlan(object):
    """Which events of a run this rank takes and which RNG streams they use.

    The run has `nevents` events of `photons_per_event` photons; photon i of event e is photon
    e * photons_per_event + i of the run and uses the RNG stream of that number, whichever rank
    it lands on."""

    def __init__(self, nevents, photons_per_event, rank=0, world_size=1):
        self.nevents, self.photons_per_event = int(nevents), int(photons_per_event)
        self.rank, self.world_size = int(rank), int(world_size)
        self.events = shard_events(nevents, rank, world_size)

    @property
    def first_stream(self):
        """Stream of this rank's first photon = base of its RNG pool."""
        return self.events.start * self.photons_per_event

    @property
    def nphotons(self):
        return len(self.events) * self.photons_per_event

    def window(self, event):
        """(first, count) of a global event's photons inside this rank's pool."""
        if event not in self.events:
            raise ValueError('event %d belongs to another rank' % event)
        return (event - self.events.start) * self.photons_per_event, self.photons_per_event


# ------------------------------------------------------------------ OR through SUM
def counter_bits(world_size):
    """Width of the per-bit counters: the smallest of 4 / 8 / 16 bits that holds world_size
    (the same rule as csrc/comm.cu history_counter_bits)."""
    return 4 if world_size <= 15 else (8 if world_size <= 255 else 16)


def pack_history(history, world_size):
    """uint32 history[C] -> uint32 [bits][C]: flag f sits in word f // (32 // bits), at bit
    (f % (32 // bits)) * bits, as a counter that starts at 0 or 1."""
    bits = counter_bits(world_size)
    per_word = 32 // bits
    h = np.asarray(history, dtype=np.uint32)
    out = np.zeros((bits, len(h)), dtype=np.uint32)
    for f in range(32):
        out[f // per_word] |= ((h >> np.uint32(f)) & np.uint32(1)) << np.uint32((f % per_word) * bits)
    return out


def unpack_history(words, world_size):
    """Summed counters -> OR of the histories."""
    bits = counter_bits(world_size)
    per_word = 32 // bits
    words = np.asarray(words, dtype=np.uint64)
    mask = np.uint64((1 << bits) - 1)
    h = np.zeros(words.shape[1], dtype=np.uint32)
    for f in range(32):
        count = (words[f // per_word] >> np.uint64((f % per_word) * bits)) & mask
        h |= (count != 0).astype(np.uint32) << np.uint32(f)
    return h


def pack_sum_buffer(q_int, history, world_size=8):
    """The SUM buffer of the collective: [q_int | packed history counters], uint32."""
    return np.concatenate([np.asarray(q_int, dtype=np.uint32), pack_history(history, world_size).ravel()])


def unpack_sum_buffer(buf, nchannels, world_size=8):
    buf = np.asarray(buf)
    q = (buf[:nchannels].astype(np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.uint32)      # wraps like atomicAdd
    h = unpack_history(buf[nchannels:].reshape(-1, nchannels), world_size)
    return q, h


def reduce_channels(time_int, q_int, history, group=None):
    """The arithmetic of cb_daq_allreduce on torch tensors, for any torch.distributed backend:
    uint32-valued int64 tensors in, reduced (time_int, q_int, history) int64 tensors out, valid on
    every rank.  Used by the CPU test tier; the product path calls the library (allreduce_daq)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    C = time_int.numel()
    packed = pack_sum_buffer(q_int.cpu().numpy(), history.cpu().numpy(), world)
    sum_buf = torch.from_numpy(packed.astype(np.int64)).to(time_int.device)
    min_buf = time_int.to(torch.int64).clone()
    if world > 1:
        w1 = dist.all_reduce(sum_buf, op=dist.ReduceOp.SUM, group=group, async_op=True)
        w2 = dist.all_reduce(min_buf, op=dist.ReduceOp.MIN, group=group, async_op=True)
        w1.wait()
        w2.wait()
    q, h = unpack_sum_buffer(sum_buf.cpu().numpy(), C, world)
    return (min_buf, torch.from_numpy(q.astype(np.int64)).to(time_int.device),
            torch.from_numpy(h.astype(np.int64)).to(time_int.device))


# ------------------------------------------------------------------ the library's communicator
def init_comm(group=None):
    """Create the library's NCCL communicator over the ranks of a torch.distributed group:
    rank 0 makes the id (cb_comm_unique_id), torch.distributed carries it, every rank joins
    (cb_comm_init).  Returns (rank, world_size); a no-op without an initialised group."""
    import ctypes as C
    from . import _lib
    lib = _lib.lib()
    try:
        import torch.distributed as dist
        live = dist.is_available() and dist.is_initialized()
    except ImportError:
        live = False
    if not live or dist.get_world_size(group) == 1:
        return 0, 1
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ident = (C.c_char * 128)()
    if rank == 0:
        _lib.check(lib.cb_comm_unique_id(ident))
    box = [bytes(ident.raw)]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ident.raw = box[0]
    _lib.check(lib.cb_comm_init(world, rank, ident))
    return rank, world


def destroy_comm():
    from . import _lib
    _lib.check(_lib.lib().cb_comm_destroy())


def allreduce_daq(gpu_daq):
    """In-place cross-GPU reduction of a GPUDaq's accumulators inside the library, float outputs
    converted behind it; returns GPUChannels valid on every rank.  Collective."""
    return gpu_daq.allreduce()


_cores_before_binding = None


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(pci_bus_id, sysfs='/sys/bus/pci/devices'):
    """Keep this process (and the threads it starts later) on the cores of the NUMA node its GPU hangs
    off, so that the page-locked event buffers it allocates from now on are local to the PCIe root the
    uploads go through.  With 8 ranks on one host every event re-uploads 130 MB per rank; measured in
    round 2 the ranks whose buffers sat on the far socket took 1 ms longer per upload.  The node comes
    from sysfs (numa_node / local_cpulist of the GPU's PCI device); nothing happens when the host
    reports no NUMA topology (-1, a VM), when the GPU's cores are not a proper subset of the cores this
    process may use, or with CHROMA_B200_NUMA=0.  Returns the cores bound to, or None."""
    import os
    global _cores_before_binding
    if os.environ.get('CHROMA_B200_NUMA', '1') == '0' or not hasattr(os, 'sched_setaffinity'):
        return None
    try:
        dev = os.path.join(sysfs, pci_bus_id.lower())
        if int(open(os.path.join(dev, 'numa_node')).read()) < 0:
            return None
        local = _parse_cpulist(open(os.path.join(dev, 'local_cpulist')).read())
        have = os.sched_getaffinity(0)
        want = local & have
        if not want or want == have:
            return None
        if _cores_before_binding is None:
            _cores_before_binding = len(have)
        os.sched_setaffinity(0, want)
        return sorted(want)
    except (OSError, ValueError):
        return None


def host_threads_should_block(local_world_size=None, cores=None):
    """Spin-waiting host threads (the CUDA default) have the lowest latency but need a core each.  Measured
    in round 2 (profiles/r02_n8_diagnostics_*.log, r02_n2_call4.log): with 4 cores per rank (8 ranks on 32
    cores, 2 ranks on 8) spinning is still the faster choice (6.7-7.4 ms per event end to end against
    8.1-8.4 ms with spin-then-block waits); only below 3 cores per rank do the three pipeline threads of a
    rank starve each other, and the waits then yield the core (cb_set_blocking_sync)."""
    import os
    if local_world_size is None:
        local_world_size = int(os.environ.get('LOCAL_WORLD_SIZE', os.environ.get('WORLD_SIZE', '1')))
    if cores is None:
        try:
            cores = _cores_before_binding or len(os.sched_getaffinity(0))    # the host's, not one NUMA node's
        except AttributeError:
            cores = os.cpu_count() or 1
    return cores / max(local_world_size, 1) < 3
