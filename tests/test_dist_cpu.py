"""World-size-2 gloo test of the multi-GPU host logic (sharding by whole clouds + the metrics gather)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_total, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pcc_b200.dist import gather_rows, shard_range, sharded_rows
    b, e = shard_range(n_total, rank, world)
    local = torch.arange(b, e, dtype=torch.float64)[:, None] * torch.tensor([1.0, 10.0, 100.0], dtype=torch.float64)
    full = gather_rows(local, n_total)
    # the scene-scale split (cfg5): every rank evaluates its slice of the query rows, one all-gather returns the whole table
    calls = []
    table = sharded_rows(lambda lo, hi: (calls.append((lo, hi)), torch.arange(lo, hi, dtype=torch.int64)[:, None].repeat(1, 4) * 3)[1],
                         n_total)
    assert calls == [(b, e)] and torch.equal(table, torch.arange(n_total, dtype=torch.int64)[:, None].repeat(1, 4) * 3)
    q.put((rank, full.tolist()))   # plain lists: a shared-memory tensor would need this process alive until it is received
    dist.destroy_process_group()


def test_shard_range_partitions():
    sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
    from pcc_b200.dist import shard_range
    for n in (0, 1, 7, 8, 10000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_gather_rows_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    import socket
    with socket.socket() as sk:          # a free port: the suite may run beside other rendezvous on this host
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    n_total, world = 7, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = torch.arange(n_total, dtype=torch.float64)[:, None] * torch.tensor([1.0, 10.0, 100.0], dtype=torch.float64)
    for r in range(world):
        assert torch.equal(torch.tensor(got[r], dtype=torch.float64), want)
