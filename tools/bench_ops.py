"""Per-op device timings at the BASELINE shapes (CUDA events, warm-up, best/median of N).  Development aid; the
judged numbers come from bench.py."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
import pcc_b200  # noqa: E402
from tools import synth  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def main():
    ops = pcc_b200.ops
    B = int(os.environ.get("B", 32))
    xyz = torch.from_numpy(synth.modelnet_like(B, 8192, seed=1)).cuda()
    y = torch.from_numpy(synth.decompressed_like(xyz.cpu().numpy())).cuda()
    start = torch.zeros(B, dtype=torch.int64, device="cuda")
    res = {}

    def rec(name, fn, pairs=None):
        best, med = timeit(fn)
        res[name] = dict(best_ms=round(best, 4), median_ms=round(med, 4))
        if pairs:
            res[name]["Gpairs_per_s"] = round(pairs / best / 1e6, 1)
        print(name, res[name], flush=True)

    rec(f"fps {B}x8192->64", lambda: ops.fps(xyz, 64, start, 1e10), B * 8192 * 64)
    cent = pcc_b200.index_points(xyz, ops.fps(xyz, 64, start, 1e10))
    rec(f"knn patch {B}x64x8192 K256", lambda: ops.knn(cent, xyz, 256, True, True, 2.0), B * 64 * 8192)
    _, _, patches = ops.knn(cent, xyz, 256, True, True, 2.0)
    patches = patches.reshape(B * 64, 256, 3)
    rec(f"knn in-patch {B*64}x256x256 K16", lambda: ops.knn(patches, patches, 16, True, True), B * 64 * 256 * 256)
    rec(f"chamfer {B}x8192x8192 (both dirs)", lambda: ops.chamfer_forward(xyz, y), 2 * B * 8192 * 8192)
    rec(f"chamfer {B}x8192x8192 no idx", lambda: ops.chamfer_forward(xyz, y, want_idx=False), 2 * B * 8192 * 8192)
    rec(f"nn1 {B}x8192x8192", lambda: ops.nn1(y, xyz), B * 8192 * 8192)
    x1, y1 = xyz[:1].contiguous(), y[:1].contiguous()
    rec("chamfer 1x8192x8192", lambda: ops.chamfer_forward(x1, y1), 2 * 8192 * 8192)
    rec("fps 1x8192->64", lambda: ops.fps(x1, 64, start[:1], 1e10), 8192 * 64)
    rec("knn patch 1x64x8192 K256", lambda: ops.knn(cent[:1].contiguous(), x1, 256, True, True, 2.0), 64 * 8192)
    sh = torch.from_numpy(synth.shapenet_like(64, 2048, seed=2)).cuda()
    rec("sfp 64x2048->512", lambda: ops.fps(sh, 512, None, ops.FLT_MAX), 64 * 2048 * 512)
    c512 = pcc_b200.index_points(sh, ops.fps(sh, 512, None, ops.FLT_MAX))
    rec("ball 64x512x2048 ns32 r.2", lambda: ops.ball_query(c512, sh, 32, 0.2), 64 * 512 * 2048)
    rec("knn 64x512x2048 K32", lambda: ops.knn(c512, sh, 32), 64 * 512 * 2048)
    feat = torch.rand(64, 2048, 128, device="cuda")
    _, bi = ops.ball_query(c512, sh, 32, 0.2)
    bi = bi.clamp(min=0)
    rec("gather 64x(512x32)x128ch", lambda: ops.gather(feat, bi))
    res["gather 64x(512x32)x128ch"]["GB_per_s"] = round(2 * 64 * 512 * 32 * 128 * 4 / res["gather 64x(512x32)x128ch"]["best_ms"] / 1e6, 1)
    # octree centre coding (8f-1): 32 clouds x 64 FPS centres, depth search + bit stream + bytes + decoded centres
    pcn, _, _, _ = ops.normalize(xyz)
    _, cxyz = ops.fps(pcn, 64, start, 1e10, return_xyz=True)
    rec(f"octree encode {B}x64 (search, bytes, stream centres)",
        lambda: ops.octree_encode(cxyz, 8192, 0.25, 0, want_bytes=True, want_stream_xyz=True))
    # whole-path rows: IPDAE compress / roundtrip per centre mode, train step (cfg2), PPPF_AE forward (cfg3)
    from pcc_b200.codec import PatchCodec
    from pcc_b200.modules import AE
    ae = AE(256, 128, 16, 7)
    ae.load_state_dict(synth.seeded_state_dict(synth.ae_shapes(128, 16, 7), 11))
    ae = ae.cuda().eval()
    for mode in ("fixed", "coded", "reference"):
        codec = PatchCodec(ae, centre_mode=mode)
        rec(f"roundtrip {B}x8192 centre_mode={mode}", lambda: codec.roundtrip(xyz, start))
        res[f"roundtrip {B}x8192 centre_mode={mode}"]["clouds_per_s"] = round(B / res[f"roundtrip {B}x8192 centre_mode={mode}"]["best_ms"] * 1e3)
    codec1 = PatchCodec(ae, centre_mode="coded")
    rec("roundtrip 1x8192 (cfg1: one cloud, compress + decompress + eval latency)", lambda: codec1.roundtrip(x1, start[:1]))
    from pcc_b200.train import Trainer
    tr = Trainer(state_dict=synth.seeded_state_dict(synth.ae_shapes(128, 16, 7), 11))
    rec(f"train step {B}x8192 K256 (cfg2: fwd + Chamfer + bwd + Adam)", lambda: tr.step(xyz, start))
    res[f"train step {B}x8192 K256 (cfg2: fwd + Chamfer + bwd + Adam)"]["clouds_per_s"] = round(
        B / res[f"train step {B}x8192 K256 (cfg2: fwd + Chamfer + bwd + Adam)"]["best_ms"] * 1e3)
    tr = Trainer(state_dict=synth.seeded_state_dict(synth.ae_shapes(128, 16, 7), 11), amp=True)
    rec(f"train step {B}x8192 K256, bf16 autocast", lambda: tr.step(xyz, start))
    del tr
    from pcc_b200 import pppf
    model = pppf.PPPF_AE(K=512, k=0, d=16, L=7)
    model.load_state_dict(synth.seeded_module_state(model, 17))
    model = model.cuda().eval()
    with torch.no_grad():
        rec("PPPF_AE forward 64x2048 (cfg3: PointNet++ SA x3 + FoldingNet)", lambda: model(sh))
        from pcc_b200 import graph as pgraph
        replay = pgraph.capture(model, sh)
        assert torch.equal(replay(sh)[0], model(sh)[0])
        rec("PPPF_AE forward 64x2048, CUDA-graph replay", lambda: replay(sh))
    res["PPPF_AE forward 64x2048 (cfg3: PointNet++ SA x3 + FoldingNet)"]["clouds_per_s"] = round(
        64 / res["PPPF_AE forward 64x2048 (cfg3: PointNet++ SA x3 + FoldingNet)"]["best_ms"] * 1e3)
    if os.environ.get("SCENE", "1") == "1":
        sc = torch.from_numpy(synth.scene_like(1_000_000, seed=3)).cuda()
        rec("fps 1x1M->7812", lambda: ops.fps(sc, 7812, start[:1], 1e10), 1e6 * 7812)
        sc_c = pcc_b200.index_points(sc, ops.fps(sc, 7812, start[:1], 1e10))
        rec("knn 1x7812x1M K256", lambda: ops.knn(sc_c, sc, 256, True, True), 7812 * 1e6)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_ops.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
