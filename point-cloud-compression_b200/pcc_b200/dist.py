"""Multi-GPU plumbing for the hot path: whole clouds are sharded across ranks (one process per GPU); the only
exchange is a gather of per-cloud eval metrics at the end of a sweep.  Works on any
torch.distributed backend (NCCL on the GPUs; gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced [begin, end) slice of n_items for `rank` (first n_items % world ranks get one extra)."""
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_rows(local, n_total, group=None):
    """All-gather row blocks of unequal length (the per-rank shard of an [n_total, C] table) into the full table,
    in shard order.  One collective: blocks are padded to the longest shard."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    longest = max(shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world))
    pad = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = []
    for r, p in enumerate(parts):
        b, e = shard_range(n_total, r, world)
        out.append(p[:e - b])
    return torch.cat(out, dim=0)


def eval_sweep(codec, clouds_host, chunk=256, rank=0, world=1, device=None):
    """cfg4: compress -> decompress -> eval over a list of clouds [n, N, 3] (host tensor), this rank's shard in chunks;
    returns the full [n, 3] metrics table (chamfer, d1_psnr, d1_mse) on every rank."""
    n = clouds_host.shape[0]
    b, e = shard_range(n, rank, world)
    rows = []
    for i in range(b, e, chunk):
        x = clouds_host[i:min(i + chunk, e)].to(device, non_blocking=True)
        start = torch.zeros(x.shape[0], dtype=torch.int64, device=device)
        rows.append(codec.roundtrip(x, start)[2])
    local = torch.cat(rows) if rows else torch.zeros((0, 3), dtype=torch.float64, device=device)
    return gather_rows(local, n)


def sharded_rows(fn, n_rows, rank=None, world=None, group=None):
    """Row-sharded evaluation of `fn(begin, end) -> tensor [end - begin, ...]` with one all-gather: every rank computes its
    contiguous slice of the n_rows rows and gets the full table back (shard order = row order).  This is the scene-scale split
    of SURVEY.md 8e: the queries of one big kNN are independent, the point cloud is replicated."""
    if rank is None or world is None:
        on = dist.is_available() and dist.is_initialized()
        rank, world = (dist.get_rank(group), dist.get_world_size(group)) if on else (0, 1)
    b, e = shard_range(n_rows, rank, world)
    return gather_rows(fn(b, e), n_rows, group=group)


def scene_patches(xyz, npoint, K, start_idx=None, group=None, return_local_nn=False, fps_idx=None):
    """compress.py:96-108 at scene scale (cfg5: one cloud of ~1M points) on `world` GPUs.  xyz [1, N, 3] is replicated on every
    rank.  FPS is 7812 dependent grid-wide arg-maxes and does not shard bit-exactly: rank 0 runs it and broadcasts the centre
    indices (the path's first collective, 8 B per centre); the kNN queries are split over the ranks against the replicated cloud
    and the index table is all-gathered (second collective, 8 B x K per centre).  Returns (fps_idx [1, npoint], knn_idx
    [1, npoint, K]) -- identical on every rank and identical to the single-GPU result -- plus, with return_local_nn, this rank's
    (begin, end, recentred patches [1, end - begin, K, 3]) so the encoder can run on the shard without another gather."""
    from . import ops
    on = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if on else 0
    world = dist.get_world_size(group) if on else 1
    N = xyz.shape[1]
    if fps_idx is None:                                                           # (given: only the kNN split runs -- timing)
        if start_idx is None:                                                     # pn_kit.py:321 (CPU RNG draw, rank 0's)
            start_idx = torch.randint(0, N, (1,), dtype=torch.long)
        if rank == 0:
            fps_idx = ops.fps(xyz, npoint, start_idx.to(xyz.device), 1e10)
        else:
            fps_idx = torch.empty((1, npoint), dtype=torch.int64, device=xyz.device)
        if on:
            dist.broadcast(fps_idx, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    centres = ops.gather(xyz, fps_idx)                                            # [1, npoint, 3]
    b, e = shard_range(npoint, rank, world)
    local = {}

    def knn_rows(b_, e_):
        q = centres[:, b_:e_].contiguous()
        _, idx, nn = ops.knn(q, xyz, K, return_nn=return_local_nn, centre_sub=True)
        local["nn"] = nn
        return idx[0]

    knn_idx = sharded_rows(knn_rows, npoint, rank, world, group)[None]
    if return_local_nn:
        return fps_idx, knn_idx, (b, e, local["nn"])
    return fps_idx, knn_idx
