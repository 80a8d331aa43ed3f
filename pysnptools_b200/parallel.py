"""Multi-GPU plumbing: one process per GPU, SNP-range sharding, one NCCL all-reduce of K.

Decode / standardize need no collective: every SNP's statistics depend on its own record only, and the SNP-major
layout makes each shard a contiguous byte range.  K = sum over shards of X_r X_r^T -- exactly the reference's
block loop (snpreader.py:651-655) run in parallel -- so the only exchange is the sum of the partial kernels.
"""
import numpy as np


def shard_range(count, rank, world):
    """Contiguous, balanced [lo, hi) of ``count`` items for ``rank`` of ``world`` (sizes differ by at most 1)."""
    base, extra = divmod(int(count), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_sum_(tensor, group=None):
    """In-place sum over the process group (NCCL on GPUs, gloo in the CPU tests). No-op without a group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


def allgather_rows(local, counts, group=None):
    """Concatenate per-rank row blocks (e.g. per-SNP statistics [m_r, 2]) in rank order."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    width = local.shape[1]
    pad = max(counts)
    buf = torch.zeros((pad, width), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


def snp_kernel_sharded(store_shard, partial_kernel_fn, n_iid, group=None, mirror_fn=None):
    """K over all shards: ``partial_kernel_fn(store_shard)`` -> (K_r lower triangle, stats_r); returns (K, stats_r)."""
    K, stats = partial_kernel_fn(store_shard)
    allreduce_sum_(K, group)
    if mirror_fn is not None:
        mirror_fn(K)
    return K, stats


def read_kernel_multi_gpu(bed, standardizer_spec=("unit",), group=None, chunk=None):
    """SnpKernel over a Bed file with the SNPs sharded over the ranks of ``group`` (every rank returns the full K)."""
    import torch
    import torch.distributed as dist
    from . import _lib, device
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lo, hi = shard_range(bed.sid_count, rank, world)
    packed = np.asarray(bed._packed_host()[lo:hi])
    store = device.PackedStore.from_host(packed, bed.iid_count)
    K, stats = device.snp_kernel(store, count_A1=bed.count_A1, standardizer=standardizer_spec, chunk=chunk, mirror=(world == 1))
    if world > 1:
        allreduce_sum_(K, group)
        _lib.check(_lib.lib.pstb_mirror_lower(K.data_ptr(), K.shape[0], K.shape[0], torch.cuda.current_stream().cuda_stream))
        counts = [shard_range(bed.sid_count, r, world)[1] - shard_range(bed.sid_count, r, world)[0] for r in range(world)]
        stats = allgather_rows(stats, counts, group)
    return K, stats
