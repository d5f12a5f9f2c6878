"""The BASELINE.json configurations beside the headline one, measured inside bench.py's run (so they sit under the driver's
clock and, under torchrun, at every N of the scaling sweep).  Each returns a small dict that bench.py puts under `configs` in its
one JSON line -- outside the timed headline region.

  cfg1  one 8192-point cloud: compress + decompress + eval latency (graph replay; and end to end from a pinned host buffer)
  cfg2  IPDAE train step, 32 clouds per rank (train.py:148-247): forward + Chamfer + backward + Adam; DistributedDataParallel
        over NCCL when world > 1 (the path's gradient all-reduce)
  cfg3  PPPF_AE forward (PointNet++ SA x3 + FoldingNet), 64 ShapeNet-shaped 2048-point clouds per rank (weak scaling)
  cfg4  eval.py's Chamfer / D1-PSNR sweep over a FIXED 10,000 clouds split over the ranks (strong scaling) + gather of the table
  cfg5  one 1M-point scene: FPS 1M -> 7812 on rank 0 + broadcast, kNN (K = 256) queries sharded over the ranks + all-gather
        (SURVEY.md 8e), and the pppe_pcd_ae encoder on the scene
"""
import time

import numpy as np
import torch

from tools import synth


class Clock:
    """CUDA-event timing of `iters` calls bracketed by a barrier + synchronize on both sides, max over ranks."""

    def __init__(self, barrier, max_over_ranks):
        self.barrier, self.max_over_ranks = barrier, max_over_ranks

    def ms(self, fn, iters, warm):
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)) / iters


def cfg1(clock, codec, dev):
    run = codec.graphed_roundtrip(1, 8192)
    x = torch.from_numpy(synth.modelnet_like(1, 8192, seed=4242)).to(dev)
    start = torch.zeros(1, dtype=torch.int64, device=dev)
    dev_ms = clock.ms(lambda: run(x, start), 50, 5)
    host = x.cpu().pin_memory()
    out = torch.empty((1, 3), dtype=torch.float64).pin_memory()
    lat = []
    for i in range(25):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out.copy_(run(host.to(dev, non_blocking=True), start)[2], non_blocking=True)
        torch.cuda.synchronize()
        if i >= 5:
            lat.append(1e3 * (time.perf_counter() - t0))
    return {"workload": "one 8192-pt cloud, K=256: compress + decompress + Chamfer/D1 (CUDA-graph replay)", "ms_per_cloud_device": dev_ms,
            "ms_per_cloud_e2e_host_to_host": float(np.median(lat)), "launches": int(run.launches)}


def cfg2(clock, dev, world, sd):
    from pcc_b200.train import Trainer
    tr = Trainer(state_dict=sd, ddp=world > 1, device=dev)
    x = torch.from_numpy(synth.modelnet_like(32, 8192, seed=77 + (torch.distributed.get_rank() if world > 1 else 0))).to(dev)
    start = torch.zeros(32, dtype=torch.int64, device=dev)
    out = {}
    eager_ms = clock.ms(lambda: out.update(tr.step(x, start)), 8, 4)
    ms = eager_ms if world > 1 else clock.ms(lambda: out.update(tr.step_graphed(x, start)), 10, 2)
    n_grad = sum(p.numel() for p in list(tr.ae.parameters()) + list(tr.prob.parameters()))
    res = {"workload": "IPDAE train step, 32 clouds x 8192 pts per rank, K=256: forward + Chamfer + backward + Adam; every "
                       "contraction of the forward and backward pass on the pcc tensor-core kernels (bf16 operands, fp32 accumulation "
                       "and weight gradients), FPS / kNN / Chamfer fwd+bwd on the pcc kernels; one CUDA-graph replay per step at 1 rank, "
                       "launched kernel by kernel under DistributedDataParallel", "ms_per_step": ms, "ms_per_step_eager": eager_ms,
           "clouds_per_s": world * 32 / ms * 1e3, "scaling": "weak", "loss": float(out["loss"]),
           "collective": (f"DistributedDataParallel gradient all-reduce over NCCL, {n_grad * 4 / 1e6:.1f} MB fp32 per step"
                          if world > 1 else "none (1 rank)")}
    del tr
    torch.cuda.empty_cache()
    return res


def cfg3(clock, dev, world):
    from pcc_b200 import graph as pgraph
    from pcc_b200 import pppf
    model = pppf.PPPF_AE(K=512, k=0, d=16, L=7)
    model.load_state_dict(synth.seeded_module_state(model, 17))
    model = model.to(dev).eval()
    sh = torch.from_numpy(synth.shapenet_like(64, 2048, seed=2)).to(dev)
    with torch.no_grad():
        eager = clock.ms(lambda: model(sh), 10, 3)
        replay = pgraph.capture(model, sh)
        ms = clock.ms(lambda: replay(sh), 20, 3)
    return {"workload": "PPPF_AE forward, 64 ShapeNet-shaped clouds x 2048 pts per rank (PointNet++ SA x3 + FoldingNet)",
            "ms_per_step": ms, "ms_per_step_eager": eager, "clouds_per_s": world * 64 / ms * 1e3, "scaling": "weak"}


def cfg4(clock, codec, dev, rank, world, total=10000, chunk=256):
    from pcc_b200 import dist as pdist
    base = torch.from_numpy(synth.modelnet_like(chunk, 8192, seed=500)).to(dev)
    noisy = torch.from_numpy(synth.decompressed_like(base.cpu().numpy(), seed=501)).to(dev)
    b, e = pdist.shard_range(total, rank, world)
    table = {}

    def chunks():
        for i in range(b, e, chunk):
            n = min(chunk, e - i)
            s = 1.0 - 1e-6 * (i // chunk)               # a different cloud set per chunk (uniform scaling keeps the structure)
            yield noisy[:n] * s, base[:n] * s

    def sweep():
        local = codec.evaluate_sweep(chunks())          # chunks alternate over two streams (build of one beside the search of the other)
        table["m"] = pdist.gather_rows(local, total)    # the sweep's only collective

    ms = clock.ms(sweep, 3, 1)
    m = table["m"]
    return {"workload": f"eval.py sweep: Chamfer + D1 PSNR of {total} clouds x 8192 pts against noisy copies, chunks of {chunk}, "
                        f"clouds split over {world} rank(s), metrics table all-gathered", "ms_per_sweep": ms,
            "clouds_per_s": total / ms * 1e3, "scaling": "strong", "rows_gathered": int(m.shape[0]),
            "mean_chamfer": float(m[:, 0].mean()), "mean_d1_psnr_db": float(m[:, 1].mean())}


def cfg5(clock, dev, rank, world, n_points=1_000_000, npoint=7812, K=256):
    from pcc_b200 import dist as pdist
    from pcc_b200 import ops, pppe
    xyz = torch.from_numpy(synth.scene_like(n_points, seed=3)).to(dev)       # replicated: every rank builds the same scene
    start = torch.zeros(1, dtype=torch.int64)
    res = {}
    fps_ms = clock.ms(lambda: ops.fps(xyz, npoint, start.to(dev), 1e10), 2, 1)   # every rank runs it here (timing only)
    out = {}
    both_ms = clock.ms(lambda: out.update(r=pdist.scene_patches(xyz, npoint, K, start)), 2, 1)
    fps_idx, knn_idx = out["r"]
    knn_ms = clock.ms(lambda: pdist.scene_patches(xyz, npoint, K, fps_idx=fps_idx), 5, 2)   # the split kNN + all-gather alone
    # single-rank check of the split: the gathered table equals this rank's own full kNN on a sample of the queries
    q = ops.gather(xyz, fps_idx)[:, :64].contiguous()
    assert torch.equal(ops.knn(q, xyz, K)[1], knn_idx[:, :64]), "sharded kNN differs from the single-GPU result"
    enc = pppe.PointNet2EncoderFull(latent_dim=256)
    enc.load_state_dict(synth.seeded_module_state(enc, 23))
    enc = enc.to(dev).eval()
    torch.manual_seed(11)
    enc_ms = clock.ms(lambda: pppe.compress(enc, xyz, latent_bins=7), 5, 2)
    res.update({"workload": f"one S3DIS-shaped scene of {n_points} pts: FPS -> {npoint} centres (rank 0 + broadcast), kNN K={K} with the "
                            f"queries split over {world} rank(s) against the replicated cloud + all-gather of the index table; "
                            "pppe_pcd_ae PointNet2EncoderFull + quantiser on the whole scene",
                "fps_ms": fps_ms, "fps_plus_sharded_knn_ms": both_ms, "knn_sharded_ms": knn_ms,
                "pppe_encoder_ms": enc_ms, "scaling": "strong (kNN queries); FPS and the encoder are single-GPU",
                "knn_table_bytes": int(knn_idx.numel() * 8)})
    return res


def run_all(dev, rank, world, barrier, max_over_ranks, codec, sd, skip=()):
    clock = Clock(barrier, max_over_ranks)
    out = {}
    for name, fn in (("cfg1", lambda: cfg1(clock, codec, dev)), ("cfg2", lambda: cfg2(clock, dev, world, sd)),
                     ("cfg3", lambda: cfg3(clock, dev, world)), ("cfg4", lambda: cfg4(clock, codec, dev, rank, world)),
                     ("cfg5", lambda: cfg5(clock, dev, rank, world))):
        if name in skip:
            continue
        t0 = time.perf_counter()
        out[name] = fn()
        torch.cuda.synchronize()
        out[name]["wall_s"] = round(time.perf_counter() - t0, 2)
    return out
