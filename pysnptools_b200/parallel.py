"""Multi-GPU plumbing: one process per GPU, SNP-range sharding, one NCCL all-reduce of K.

Decode / standardize need no collective: every SNP's statistics depend on its own record only, and the SNP-major
layout makes each shard a contiguous byte range.  K = sum over shards of X_r X_r^T -- exactly the reference's
block loop (snpreader.py:651-655) run in parallel -- so the only exchange is the sum of the partial kernels.
"""
import numpy as np


def bind_to_gpu_numa_node(device=None):
    """One process per GPU on a multi-socket host: bind this process' calling thread (and the threads created after it -- the
    library's copy workers, NCCL's proxies) to the CPUs of the NUMA node the GPU hangs off and prefer that node's memory
    (``pstb_numa_bind``), so the page-locked buffers of the host-buffer entry points are allocated next to the GPU's PCIe root.
    Call it before the first read / ``init_process_group``.  Returns the node, or -1 when the host exposes none (nothing changed)."""
    import torch
    from . import _lib
    dev = torch.cuda.current_device() if device is None else int(device)
    node = int(_lib.lib.pstb_numa_bind(dev))
    if node == -2:
        raise RuntimeError(_lib.last_error())
    return node


def shard_range(count, rank, world):
    """Contiguous, balanced [lo, hi) of ``count`` items for ``rank`` of ``world`` (sizes differ by at most 1)."""
    base, extra = divmod(int(count), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def assign_pieces(sizes, world):
    """DistributedBed pieces -> ranks: longest piece first onto the least loaded rank (SNP counts as weights).  Every piece goes
    to exactly one rank; returns ``world`` ascending lists of piece numbers.  Pieces are SNP ranges, i.e. the per-GPU shard unit
    of the kinship path, so a rank reads only its own piece files."""
    loads = [0] * int(world)
    owned = [[] for _ in range(int(world))]
    for k in sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i)):
        r = min(range(int(world)), key=lambda q: (loads[q], q))
        owned[r].append(k)
        loads[r] += int(sizes[k])
    return [sorted(o) for o in owned]


def allreduce_sum_(tensor, group=None):
    """In-place sum over the process group (NCCL on GPUs, gloo in the CPU tests). No-op without a group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


def allreduce_tiles_and_expand(tiles, n_iid, K, group=None, slices=4, u=None):
    """Sum the compact lower-triangular tiles ``[count, 256, 256]`` over the ranks and expand them into the square ``K``: ONE logical
    reduction issued in ``slices`` asynchronous NCCL calls, so that ``pstb_kernel_from_tiles_range`` of a reduced slice runs while the
    next slice is still on the wire (the expansion moves 3 x the bytes of the reduction through HBM: 6 of 24 ms at 8 GPUs when serial).
    ``u``: this rank's deferred rank-one vector (float64 [n]); it is all-reduced as well and added during the expansion, which saves
    every rank a read-modify-write sweep over its 5 GB of tiles."""
    import torch
    import torch.distributed as dist
    from . import _lib
    count = int(tiles.shape[0])
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    slices = max(1, min(int(slices), count)) if count else 1
    edges = [count * s // slices for s in range(slices + 1)]
    works = []
    u_work = dist.all_reduce(u, op=dist.ReduceOp.SUM, group=group, async_op=True) if (multi and u is not None) else None
    for s in range(slices):
        t0, t1 = edges[s], edges[s + 1]
        works.append(dist.all_reduce(tiles[t0:t1], op=dist.ReduceOp.SUM, group=group, async_op=True) if (multi and t1 > t0) else None)
    stream = torch.cuda.current_stream().cuda_stream
    if u_work is not None:
        u_work.wait()
    for s in range(slices):
        t0, t1 = edges[s], edges[s + 1]
        if t1 <= t0:
            continue
        if works[s] is not None:
            works[s].wait()                         # the current stream waits for this slice only
        _lib.check(_lib.lib.pstb_kernel_from_tiles_range(tiles.data_ptr(), int(n_iid), 0, 1, t0, t1, K.data_ptr(),
                                                         u.data_ptr() if u is not None else None, stream))
    return K


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


def snp_kernel_sharded_overlapped(store, n_iid, total_sid, group=None, standardizer_spec=("unit",), count_A1=False, chunk=None,
                                  bands=8, reserve_sms=0, tail_chunks=4, tiles=None, K=None):
    """SNP-sharded SnpKernel with the NCCL reduction OVERLAPPED: this rank's partial K accumulates in compact
    lower-triangular tiles.  K is final only after the last SNP chunk, so the last ``tail_chunks`` chunks are multiplied BAND-major
    instead of chunk-major (``pstb_snp_kernel_tiles_band``: their operand planes are built once, into one workspace each, then every
    band of tiles takes all of them in turn): as soon as a band is final it is all-reduced on a side stream and expanded into the
    square K while the later bands still multiply (``reserve_sms`` SMs are left to the collective).  Only the last band's reduction
    (1 / bands of the triangle) stays exposed.  Returns ``(K, stats_local)``: float32 CUDA tensor [n, n] (both triangles) and this
    rank's float64 statistics [m_local, 2]."""
    import os
    import torch
    import torch.distributed as dist
    from . import _lib, device
    lib, check = _lib.lib, _lib.check
    dev = store.device
    n, m_local = int(n_iid), store.sid_count
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    chunk = int(chunk or device.default_kernel_chunk(n, m_local))
    low_term = device.low_term_for(total_sid, n, standardizer_spec)
    coords = device.kernel_tile_coords(n)
    ntiles = len(coords)
    trace = os.environ.get("PSTB_OVERLAP_TRACE") and (not dist.is_initialized() or dist.get_rank(group) == 0)
    with torch.cuda.device(dev):
        if tiles is None:
            tiles = torch.empty((ntiles, 256, 256), dtype=torch.float32, device=dev)
        if K is None:
            K = torch.empty((n, n), dtype=torch.float32, device=dev)
        stats = torch.empty((m_local, 2), dtype=torch.float64, device=dev)
        if m_local == 0:
            tiles.zero_()
        nchunks = (m_local + chunk - 1) // chunk
        ntail = max(1, min(int(tail_chunks), nchunks)) if nchunks else 0
        head = (nchunks - ntail) * chunk if nchunks else 0
        main = torch.cuda.current_stream()
        t_ev = [torch.cuda.Event(enable_timing=True)] if trace else None
        if trace:
            t_ev[0].record(main)
        u_total = torch.zeros(n, dtype=torch.float64, device=dev)        # deferred rank-one vectors of all calls (added during the expansion)
        if head > 0:
            _t, _c, st_head, u_head = device.snp_kernel_tiles(store, None, slice(0, head), count_A1=count_A1, standardizer=standardizer_spec,
                                                              chunk=chunk, tiles=tiles, accumulate=False, low_term=low_term, defer_rank1=True)
            stats[:head] = st_head
            u_total += u_head
        bands = max(1, min(int(bands), ntiles))
        edges = [ntiles * b // bands for b in range(bands + 1)]
        # high priority: when SMs free up at the end of a band, the collective's CTAs are placed before the next SYRK launch's, so the
        # persistent CTA pairs (statically assigned tiles) never find their SMs taken half-way through a launch
        comm = torch.cuda.Stream(device=dev, priority=-1)
        comm.wait_stream(main)
        mode, a, b = device._mode_args(standardizer_spec)
        wbytes = int(lib.pstb_kernel_workspace_bytes(n, chunk))
        works = [torch.empty(wbytes, dtype=torch.uint8, device=dev) for _ in range(ntail)]
        iid_ax = _lib.Axis(None, 0, 1, n)
        ranges = [(head + c * chunk, min(m_local, head + (c + 1) * chunk)) for c in range(ntail)]
        reserve = int(reserve_sms) if world > 1 else 0

        def band_call(c, t0, t1, flags):
            lo, hi = ranges[c]
            check(lib.pstb_snp_kernel_tiles_band(store.tensor.data_ptr(), store.ld, store.iid_count, store.sid_count, iid_ax,
                                                 _lib.Axis(None, lo, 1, hi - lo), int(bool(count_A1)), mode, a, b, 0, stats[lo:hi].data_ptr(),
                                                 tiles.data_ptr(), 0, 1, int(head > 0 or c > 0), works[c].data_ptr(), wbytes, chunk,
                                                 device._LOW_TERM[low_term], t0, t1, flags, reserve, main.cuda_stream))
        for c in range(ntail):
            band_call(c, 0, 0, 1 | 2)                               # statistics + operand planes of the tail chunks, no tiles yet
            u_total += device.workspace_rank1(works[c], n, chunk)
        with torch.cuda.stream(comm):
            comm.wait_stream(main)
            if world > 1:
                dist.all_reduce(u_total, op=dist.ReduceOp.SUM, group=group)
        marks = []
        for bnd in range(bands):
            t0, t1 = edges[bnd], edges[bnd + 1]
            if t1 <= t0:
                continue
            for c in range(ntail):
                band_call(c, t0, t1, 2)
            ev = torch.cuda.Event(enable_timing=bool(trace))
            ev.record(main)
            with torch.cuda.stream(comm):
                comm.wait_event(ev)
                if world > 1:
                    dist.all_reduce(tiles[t0:t1], op=dist.ReduceOp.SUM, group=group)
                check(lib.pstb_kernel_from_tiles_range(tiles.data_ptr(), n, 0, 1, t0, t1, K.data_ptr(), u_total.data_ptr(), comm.cuda_stream))
                if trace:
                    e2 = torch.cuda.Event(enable_timing=True)
                    e2.record(comm)
                    marks.append((ev, e2))
        main.wait_stream(comm)
        tiles.record_stream(comm)
        K.record_stream(comm)
        u_total.record_stream(comm)
        for w in works:
            w.record_stream(main)
        if trace:
            torch.cuda.synchronize()
            print("[overlap] head %d SNPs, %d tail chunks, %d bands, %d SMs reserved; ms since start: " % (head, ntail, bands, reserve)
                  + "  ".join("band%d compute %.1f reduced %.1f" % (i, t_ev[0].elapsed_time(a_), t_ev[0].elapsed_time(b_)) for i, (a_, b_) in enumerate(marks)), flush=True)
    return K, stats


def distributed_bed_partial_kernel(dbed, rank, world, standardizer_spec=("unit",), chunk=None):
    """This rank's share of ``K`` for a :class:`DistributedBed`: the lower triangle accumulated over the pieces
    ``assign_pieces`` gives it (each piece file is read by one rank only).  Returns ``(K_r, stats_r, sid_positions_r)``:
    float32 CUDA tensor [n, n] (not mirrored), float64 CUDA tensor [m_r, 2], and the positions of those SNPs in ``dbed.sid``."""
    import torch
    from . import device
    dbed._run_once()
    sizes = [p.sid_count for p in dbed._pieces]
    mine = assign_pieces(sizes, world)[rank]
    n = dbed.iid_count
    K = torch.zeros((n, n), dtype=torch.float32, device="cuda")
    stats, where = [], []
    for k in mine:
        piece = dbed._pieces[k]
        store = device.PackedStore.from_host(np.asarray(piece._packed_host()), n)
        K, st = device.snp_kernel(store, count_A1=piece.count_A1, standardizer=standardizer_spec, chunk=chunk, K=K, accumulate=True, mirror=False,
                                  low_term=device.low_term_for(dbed.sid_count, n, standardizer_spec))
        stats.append(st)
        where.append(np.arange(dbed._starts[k], dbed._starts[k + 1], dtype=np.int64))
    stats = torch.cat(stats) if stats else torch.zeros((0, 2), dtype=torch.float64, device="cuda")
    return K, stats, (np.concatenate(where) if where else np.zeros(0, dtype=np.int64))


def read_kernel_multi_gpu(bed, standardizer_spec=("unit",), group=None, chunk=None):
    """SnpKernel over a Bed file (SNP ranges) or a DistributedBed (whole pieces) sharded over the ranks of ``group``; one
    all-reduce of the partial kernels; every rank returns the full K and the statistics in SNP order."""
    import torch
    import torch.distributed as dist
    from . import _lib, device
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if hasattr(bed, "_read_kernel_pieces"):                                  # DistributedBed
        K, stats, where = distributed_bed_partial_kernel(bed, rank, world, standardizer_spec, chunk)
        if world > 1:
            allreduce_sum_(K, group)
            owned = assign_pieces([p.sid_count for p in bed._pieces], world)
            counts = [int(sum(bed._pieces[k].sid_count for k in o)) for o in owned]
            stats = allgather_rows(stats, counts, group)
            where = np.concatenate([np.arange(bed._starts[k], bed._starts[k + 1], dtype=np.int64) for o in owned for k in o])
        _lib.check(_lib.lib.pstb_mirror_lower(K.data_ptr(), K.shape[0], K.shape[0], torch.cuda.current_stream().cuda_stream))
        ordered = torch.empty_like(stats)
        ordered[torch.as_tensor(where, device=stats.device)] = stats         # rank-major -> SNP order
        return K, ordered
    lo, hi = shard_range(bed.sid_count, rank, world)
    packed = np.asarray(bed._packed_host()[lo:hi])
    store = device.PackedStore.from_host(packed, bed.iid_count)
    if world == 1:
        K, stats = device.snp_kernel(store, count_A1=bed.count_A1, standardizer=standardizer_spec, chunk=chunk, mirror=True)
    else:
        # partial kernel in compact lower-triangular tile storage: the all-reduce moves half the bytes of the square matrix, band by band
        # under the multiplication of the last SNP chunks (the low-term mode follows the SNP count of the whole kernel)
        K, stats = snp_kernel_sharded_overlapped(store, bed.iid_count, bed.sid_count, group, standardizer_spec, bed.count_A1, chunk)
        counts = [shard_range(bed.sid_count, r, world)[1] - shard_range(bed.sid_count, r, world)[0] for r in range(world)]
        stats = allgather_rows(stats, counts, group)
    return K, stats
