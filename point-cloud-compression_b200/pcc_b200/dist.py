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
