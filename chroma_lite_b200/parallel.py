"""Multi-GPU plumbing: one process per GPU, photon banks partitioned, geometry
replicated, per-channel DAQ accumulators combined with ONE reduction.

The reference has no multi-GPU path; its atomics on
per-channel arrays (daq.cu:73-75) are what becomes a collective once photons
are sharded:  earliest_time_int -> MIN, channel_q_int -> SUM, channel_history
-> bitwise OR.  NCCL has no OR, so the history word travels as 16 per-bit
counters inside the SUM buffer (the kernel only ever sets 16 bits).  Both
reductions go out in a single coalesced group over NVLink.
"""
import numpy as np

HISTORY_BITS = 16


def shard_range(n, rank, world_size):
    """Contiguous [start, end) of `n` items owned by `rank` (balanced to +-1)."""
    base, rem = divmod(int(n), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_events(nevents, rank, world_size):
    """Whole events per rank so evidx and per-event DAQ stay local."""
    return range(*shard_range(nevents, rank, world_size))


def pack_sum_buffer(q_int, history):
    """uint32 q_int[C], history[C] -> int64 [C*(1+16)] additive buffer."""
    q = np.asarray(q_int, dtype=np.int64)
    h = np.asarray(history, dtype=np.uint32)
    bits = ((h[:, None] >> np.arange(HISTORY_BITS, dtype=np.uint32)[None, :]) & 1).astype(np.int64)
    return np.concatenate([q, bits.ravel()])


def unpack_sum_buffer(buf, nchannels):
    buf = np.asarray(buf)
    q = (buf[:nchannels] & 0xFFFFFFFF).astype(np.uint32)          # uint32 wrap-around like atomicAdd
    bits = buf[nchannels:].reshape(nchannels, HISTORY_BITS) > 0
    h = (bits.astype(np.uint32) << np.arange(HISTORY_BITS, dtype=np.uint32)[None, :]).sum(axis=1).astype(np.uint32)
    return q, h


def reduce_channels(time_int, q_int, history, group=None, dst=None):
    """Combine per-rank DAQ accumulators across the process group.

    time_int/q_int/history: torch tensors (int64 views of the uint32 device
    arrays, on the device for NCCL or on the CPU for gloo).  Returns the reduced
    (time_int, q_int, history) as torch int64 tensors valid on every rank
    (dst=None -> all_reduce) or on `dst` only (reduce)."""
    import torch
    import torch.distributed as dist
    C = time_int.numel()
    shifts = torch.arange(HISTORY_BITS, device=history.device, dtype=torch.int64)
    bits = ((history.to(torch.int64)[:, None] >> shifts[None, :]) & 1).reshape(-1)
    sum_buf = torch.cat([q_int.to(torch.int64), bits])
    min_buf = time_int.to(torch.int64).clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if dst is None:
            w1 = dist.all_reduce(sum_buf, op=dist.ReduceOp.SUM, group=group, async_op=True)
            w2 = dist.all_reduce(min_buf, op=dist.ReduceOp.MIN, group=group, async_op=True)
        else:
            w1 = dist.reduce(sum_buf, dst=dst, op=dist.ReduceOp.SUM, group=group, async_op=True)
            w2 = dist.reduce(min_buf, dst=dst, op=dist.ReduceOp.MIN, group=group, async_op=True)
        w1.wait()
        w2.wait()
    q = sum_buf[:C] & 0xFFFFFFFF
    h = ((sum_buf[C:].reshape(C, HISTORY_BITS) > 0).to(torch.int64) << shifts[None, :]).sum(dim=1)
    return min_buf, q, h


def reduce_daq(gpu_daq, group=None, dst=None):
    """In-place cross-GPU reduction of a GPUDaq's integer accumulators followed
    by the fused finaliser; returns GPUChannels (meaningful on dst / all ranks)."""
    import torch
    n = gpu_daq.earliest_time_int_gpu.size
    dev = torch.device('cuda', torch.cuda.current_device())
    as_i64 = lambda a: torch.as_tensor(a.view(np.int32), device=dev).to(torch.int64) & 0xFFFFFFFF
    t, q, h = reduce_channels(as_i64(gpu_daq.earliest_time_int_gpu), as_i64(gpu_daq.channel_q_int_gpu),
                              as_i64(gpu_daq.channel_history_gpu), group=group, dst=dst)
    for arr, val in ((gpu_daq.earliest_time_int_gpu, t), (gpu_daq.channel_q_int_gpu, q),
                     (gpu_daq.channel_history_gpu, h)):
        host = val.to('cpu').numpy().astype(np.uint32)
        arr.set(host)
    assert n == len(host)
    return gpu_daq.finalize_reduced()
