"""Mint the golden vectors under tests/golden/ (run HERE, where /root/reference exists):

    python tests/golden/make_golden.py

* ref_fps_gather.npz  -- outputs of the REFERENCE's own pn_kit.farthest_point_sample_batch / index_points
                         (imported from /root/reference), including the CPU-RNG start indices it drew.
* ref_octree.npz      -- outputs of the REFERENCE's own octree centre coder (pn_kit.encode_sampled_np, decode_sampled_np,
                         binary_array_to_byte_array, octree_np.encode, getDecodeFromPc) on seeded FPS centres.
* ref_eval.npz        -- eval.py's calc_uc run from the reference's own source (extracted with ast), and the oracle's
                         p2plane PSNR (Open3D absent: unpinned).
* ref_entropy.npz     -- outputs of the REFERENCE's own pn_kit.pmf_to_cdf / estimate_bits_from_pmf on seeded PMFs.
* pppe_modules.npz    -- outputs of the REFERENCE's own pppe_pcd_ae.PointNet2EncoderFull on CPU (kNN served by the oracle).
* p3d_ops.npz         -- outputs of the oracle restatement of the PyTorch3D ops (knn_points, ball_query,
                         sample_farthest_points, chamfer_distance) on seeded inputs, cross-checked here against
                         an independent torch brute-force statement before being written.  PyTorch3D itself
                         cannot be installed in this image, so these are "parity unpinned" w.r.t. the package.
Inputs are regenerated from seeds by the tests (tools/synth.py); the files keep inputs too, so a drift in
the generator is caught.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402
from oracle import ref_loader  # noqa: E402
from tools import synth  # noqa: E402


def ref_fps_gather():
    pn = ref_loader.load("pn_kit")
    out = {}
    cases = {
        "modelnet": (synth.modelnet_like(2, 8192, seed=11), 64),
        "grid": (synth.grid_quantised(3, 1000, depth=3, seed=12), 37),
        "perm": (synth.uniform_cube(2, 257, seed=13), 257),
        "tiny": (synth.uniform_cube(4, 16, seed=14), 5),
    }
    for name, (xyz, npoint) in cases.items():
        torch.manual_seed(11)  # the reference scripts seed 11 (train.py:18-20)
        state = torch.get_rng_state()
        t = torch.from_numpy(xyz)
        idx = pn.farthest_point_sample_batch(t, npoint)
        torch.set_rng_state(state)
        start = torch.randint(0, xyz.shape[1], (xyz.shape[0],), dtype=torch.long)
        assert torch.equal(idx[:, 0], start)
        g = pn.index_points(t, idx)
        out[f"{name}_xyz"] = xyz
        out[f"{name}_npoint"] = np.int64(npoint)
        out[f"{name}_start"] = start.numpy()
        out[f"{name}_idx"] = idx.numpy()
        out[f"{name}_gather"] = g.numpy()
        # [B,S,K] form of index_points
        idx3 = torch.from_numpy(np.random.default_rng(5).integers(0, xyz.shape[1], (xyz.shape[0], 7, 3)))
        out[f"{name}_idx3"] = idx3.numpy()
        out[f"{name}_gather3"] = pn.index_points(t, idx3).numpy()
        assert np.array_equal(orc.fps(xyz, npoint, start.numpy(), 1e10), idx.numpy()), name
        assert np.array_equal(orc.gather(xyz, idx.numpy()), g.numpy()), name
    np.savez_compressed(os.path.join(HERE, "ref_fps_gather.npz"), **out)
    print("ref_fps_gather.npz:", {k: v.shape for k, v in out.items() if k.endswith("_idx")})


def _torch_knn(q, p, K):
    D = ((torch.from_numpy(q)[:, :, None, :] - torch.from_numpy(p)[:, None, :, :]) ** 2).sum(-1)
    ds, js = torch.sort(D, dim=2, stable=True)
    return ds[:, :, :K].numpy(), js[:, :, :K].numpy()


def p3d_ops():
    out = {}
    # kNN: smooth + tie-heavy + duplicated queries (the reference's <= 8 distinct centres, appendix B-2)
    p = synth.modelnet_like(2, 2048, seed=21)
    q = orc.gather(p, orc.fps(p, 16, np.zeros(2, np.int64), 1e10))
    qg = ((np.floor(q * 2) + 0.5) / 2).astype(np.float32)  # depth-1 octant centres, many duplicates
    pg = synth.grid_quantised(2, 700, depth=3, seed=22)
    for name, (a, b, K) in {"knn_patch": (q, p, 256), "knn_dupq": (qg, p, 64), "knn_ties": (pg[:, :50], pg, 33),
                            "knn_self": (p[:, :256], p[:, :256], 16), "knn_kgtn": (q, p[:, :10], 16)}.items():
        d, i, nn = orc.knn_points(a, b, K, True)
        if b.shape[1] >= K:
            td, ti = _torch_knn(a, b, K)
            assert np.array_equal(td, d) and np.array_equal(ti, i), name
        out[f"{name}_q"], out[f"{name}_p"], out[f"{name}_K"] = a, b, np.int64(K)
        out[f"{name}_d"], out[f"{name}_i"], out[f"{name}_nn"] = d, i, nn
    # ball query incl. the exact boundary d2 == r2 (grid points at distance exactly 0.25)
    for name, (a, b, K, r) in {"ball_sa1": (q, p, 32, 0.2), "ball_edge": (pg[:, :40], pg, 16, 0.25),
                               "ball_few": (q, p, 64, 0.05)}.items():
        d, i = orc.ball_query(a, b, K, r)
        D = ((torch.from_numpy(a)[:, :, None, :] - torch.from_numpy(b)[:, None, :, :]) ** 2).sum(-1).numpy()
        r2 = np.float32(r) * np.float32(r)
        for bb in range(a.shape[0]):
            for qq in range(a.shape[1]):
                hits = np.nonzero(D[bb, qq] < r2)[0][:K]
                assert np.array_equal(i[bb, qq, :len(hits)], hits) and np.all(i[bb, qq, len(hits):] == -1), name
        out[f"{name}_q"], out[f"{name}_p"], out[f"{name}_K"], out[f"{name}_r"] = a, b, np.int64(K), np.float32(r)
        out[f"{name}_d"], out[f"{name}_i"] = d, i
    # sample_farthest_points (start 0, FLT_MAX init, -1 padding when K > N)
    for name, (a, K) in {"sfp": (p, 128), "sfp_pad": (p[:, :20], 32), "sfp_ties": (pg, 64)}.items():
        pts, idx = orc.sample_farthest_points(a, K)
        out[f"{name}_x"], out[f"{name}_K"], out[f"{name}_idx"], out[f"{name}_pts"] = a, np.int64(K), idx, pts
    # chamfer
    x = synth.modelnet_like(2, 1024, seed=23)
    y = synth.decompressed_like(x, seed=24)[:, :900]
    loss, pc, dx, ix, dy, iy = orc.chamfer(x, y)
    D = ((torch.from_numpy(x)[:, :, None, :] - torch.from_numpy(y)[:, None, :, :]) ** 2).sum(-1)
    tdx, tix = D.min(2)
    tdy, tiy = D.min(1)
    assert np.array_equal(tdx.numpy(), dx) and np.array_equal(tdy.numpy(), dy)
    assert np.array_equal(tix.numpy(), ix) and np.array_equal(tiy.numpy(), iy)
    tl = (tdx.sum(1) / 1024 + tdy.sum(1) / 900).sum() / 2
    assert abs(tl.item() - loss) <= 1e-5 * abs(loss)
    gx, gy = orc.chamfer_bwd(x, y, ix, iy, 1.0)
    out.update(cham_x=x, cham_y=y, cham_loss=np.float64(loss), cham_pc=pc, cham_dx=dx, cham_ix=ix, cham_dy=dy,
               cham_iy=iy, cham_gx=gx, cham_gy=gy)
    psnr, mse = orc.d1_psnr(x[0], y[0])
    out.update(d1_psnr=np.float64(psnr), d1_mse=np.float64(mse))
    np.savez_compressed(os.path.join(HERE, "p3d_ops.npz"), **out)
    print("p3d_ops.npz written:", len(out), "arrays")




def pppf_modules():
    """Outputs of the REFERENCE's PPPF_AE (PointnetSAModule x3 + FoldingNet) in eval mode on CPU, PyTorch3D ops served by
    the oracle, with the seeded state of tools/synth.seeded_module_state (regenerated, not stored, by the tests)."""
    ref = ref_loader.load("PPPF_AE")
    model = ref.PPPF_AE(K=512, k=0, d=16, L=7)
    model.load_state_dict(synth.seeded_module_state(model, 17))
    model.eval()
    x = torch.from_numpy(synth.shapenet_like(2, 2048, seed=51))
    with torch.no_grad():
        xyz1, f1 = model.encoder.sa1(x, None)
        xyz2, f2 = model.encoder.sa2(xyz1, f1)
        xyz3, f3 = model.encoder.sa3(xyz2, f2)
        recon, latent, lq = model(x)
    np.savez_compressed(os.path.join(HERE, "pppf_modules.npz"), x=x.numpy(), xyz1=xyz1.numpy(), f1=f1.numpy(), xyz2=xyz2.numpy(),
                        f2=f2.numpy(), xyz3=xyz3.numpy(), f3=f3.numpy(), recon=recon.numpy(), latent=latent.numpy(), lq=lq.numpy())
    print("pppf_modules.npz:", {k: tuple(v.shape) for k, v in dict(f1=f1, f2=f2, f3=f3, recon=recon, latent=latent).items()})


def pppe_modules():
    """Outputs of the REFERENCE's pppe_pcd_ae.PointNet2EncoderFull (MSG + 2 SS set-abstraction levels, global head) in eval mode
    on CPU, kNN served by the oracle, FPS = the reference's own pn_kit function with its CPU-RNG start draws (seed 11 first)."""
    ref = ref_loader.load("pppe_pcd_ae")
    model = ref.PointNet2EncoderFull(latent_dim=256)
    model.load_state_dict(synth.seeded_module_state(model, 23))
    model.eval()
    x = torch.from_numpy(synth.modelnet_like(2, 2048, seed=71))
    with torch.no_grad():
        torch.manual_seed(11)
        latent, pooled = model(x)
        torch.manual_seed(11)
        xyz1, f1 = model.sa_modules[0](x, None)
        xyz2, f2 = model.sa_modules[1](xyz1, f1)
    np.savez_compressed(os.path.join(HERE, "pppe_modules.npz"), x=x.numpy(), latent=latent.numpy(), pooled=pooled.numpy(),
                        xyz1=xyz1.numpy(), f1=f1.numpy(), xyz2=xyz2.numpy(), f2=f2.numpy())
    print("pppe_modules.npz:", {k: tuple(v.shape) for k, v in dict(latent=latent, pooled=pooled, f1=f1, f2=f2).items()})


def ae_modules():
    """Outputs of the REFERENCE's AE.AE (SetAbstraction + PointNet encoder, inv_pool + MLP decoder) on CPU."""
    ref = ref_loader.load("AE")
    model = ref.AE(K=256, k=128, d=16, L=7)
    model.load_state_dict(synth.seeded_state_dict(synth.ae_shapes(128, 16, 7), 11))
    model.eval()
    x = torch.from_numpy(synth.uniform_cube(6, 256, seed=52)) - 0.5
    with torch.no_grad():
        new_xyz, latent, lq = model(x)
        _, feat = model.sa(x.transpose(2, 1))
    np.savez_compressed(os.path.join(HERE, "ae_modules.npz"), x=x.numpy(), new_xyz=new_xyz.numpy(), latent=latent.numpy(),
                        lq=lq.numpy(), sa_feat=feat.numpy())
    print("ae_modules.npz:", tuple(new_xyz.shape), tuple(latent.shape))


def octree_cases():
    """name -> (centres [B,S,3] in (0,1), N, min_bpp): the inputs of the octree goldens (regenerated by the tests)."""
    def centres(n_clouds, n_points, S, seed):
        p = synth.modelnet_like(n_clouds, n_points, seed=seed)
        mn, mx = p.min(axis=1, keepdims=True), p.max(axis=1, keepdims=True)  # pn_kit.normalize's range [0.01, 0.99]
        p = ((p - (mx + mn) / 2) * np.float32(0.99) / (mx - mn).max(axis=2, keepdims=True) + np.float32(0.5)).astype(np.float32)
        return orc.gather(p, orc.fps(p, S, np.zeros(n_clouds, np.int64), 1e10))
    c64 = centres(5, 8192, 64, 61)
    dup = c64[:2].copy()
    dup[:, 40:] = dup[:, :24]                      # two centres in one cell at every depth: the search runs to depth 16
    grid = synth.grid_quantised(2, 64, depth=3, seed=62)
    return {
        "k256": (c64, 8192, 0.25),                 # the headline setting: K=256 -> OCTREE_BPP_DICT[256] (pn_kit.py:17-23)
        "k128": (centres(2, 8192, 128, 63), 8192, 0.5),
        "k1024": (centres(2, 8192, 16, 64), 8192, 0.07),
        "s300": (centres(1, 4096, 300, 65), 4096, 1.0),   # more than one chunk of 256 sorted keys per CTA
        "dup": (dup, 8192, 0.25),
        "grid": (grid, 8192, 0.01),                # cell-centre inputs (ties at the snapping boundaries), tiny bpp target
        "one": (c64[:3, :1].copy(), 8192, 0.0001),
        "edge": (np.array([[[0.0, 0.0, 0.0], [0.99999994, 0.99999994, 0.99999994], [0.5, 0.5, 0.5], [0.49999997, 0.5, 0.25]]],
                          dtype=np.float32), 16, 0.25),
    }


def ref_octree():
    """Outputs of the REFERENCE's own octree centre coder (pn_kit.encode_sampled_np / decode_sampled_np /
    binary_array_to_byte_array, octree_np.encode / getDecodeFromPc) imported from /root/reference."""
    pn = ref_loader.load("pn_kit")
    on = ref_loader.load("octree_np")
    out = {}
    for name, (c, N, min_bpp) in octree_cases().items():
        codes, codebits = pn.encode_sampled_np(c, scale=1, N=N, min_bpp=min_bpp)
        rec = pn.decode_sampled_np(codes, scale=1)
        ocodes, obits, depths = orc.encode_sampled_np(c, 1, N, min_bpp)
        assert codebits == obits, name
        nb = np.array([len(x) for x in codes], np.int32)
        bits = np.zeros((len(codes), nb.max()), np.uint8)
        by = np.zeros((len(codes), (nb.max() + 7) // 8), np.uint8)
        for b, code in enumerate(codes):
            assert np.array_equal(code, ocodes[b]), name
            assert np.array_equal(code, on.encode(c[b], 1, depths[b])), name   # the returned code is the depth-`depths[b]` code
            assert np.array_equal(rec[b], orc.octree_decode_ref(code)), name
            bs = bytes(pn.binary_array_to_byte_array(code))
            assert bs == orc.bits_to_bytes(code).tobytes(), name
            bits[b, :nb[b]] = code
            by[b, :len(bs)] = np.frombuffer(bs, np.uint8)
            u = on.getDecodeFromPc(c[b], 1, depths[b])
            snapped, uu = orc.octree_quantise(c[b], 1, depths[b])
            assert np.array_equal(u, uu), name
            out[f"{name}_uniq{b}"] = u
        out[f"{name}_c"], out[f"{name}_N"], out[f"{name}_min_bpp"] = c, np.int64(N), np.float64(min_bpp)
        out[f"{name}_bits"], out[f"{name}_nbits"], out[f"{name}_depth"] = bits, nb, np.array(depths, np.int32)
        out[f"{name}_bytes"], out[f"{name}_rec"] = by, rec.astype(np.float32)
    # fixed-depth octree_np.encode
    c = octree_cases()["k256"][0][0]
    for d in (1, 2, 5, 9):
        out[f"fixed_d{d}"] = on.encode(c, 1, d)
    np.savez_compressed(os.path.join(HERE, "ref_octree.npz"), **out)
    print("ref_octree.npz:", {k: (out[k].tolist()) for k in out if k.endswith("_depth")})


def ref_eval():
    """eval.py metrics.  calc_uc is the REFERENCE's own function: eval.py cannot be imported (it parses argv and globs files at
    import), so the function definition is extracted from its source with `ast` and executed with the stubbed knn_points.
    The p2plane numbers are the oracle's (Open3D is not installed: parity unpinned)."""
    import ast
    ref_loader.install_stubs()
    tree = ast.parse(open(os.path.join(ref_loader.REFERENCE_ROOT, "eval.py")).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "calc_uc"][0]
    ns = {"np": np, "torch": torch, "knn_points": sys.modules["pytorch3d.ops.knn"].knn_points}
    exec(compile(ast.Module([fn], []), "eval.py", "exec"), ns)
    x = synth.modelnet_like(3, 8192, seed=81)
    y = synth.decompressed_like(x, seed=82)
    uc = np.array([ns["calc_uc"](x[b], y[b]) for b in range(3)], np.float64)
    assert np.allclose(uc, [orc.calc_uc(x[b], y[b]) for b in range(3)], rtol=1e-4)   # cdist's matmul path: 6th digit
    pp = np.array([orc.p2plane_psnr(x[b], y[b]) for b in range(3)], np.float64)
    np.savez_compressed(os.path.join(HERE, "ref_eval.npz"), uc=uc, p2plane=pp)
    print("ref_eval.npz: uc", uc, "p2plane", pp[:, 0])


def ref_entropy():
    """pn_kit.pmf_to_cdf (the REFERENCE's own function, pn_kit.py:452-461) on seeded PMFs, and pn_kit.estimate_bits_from_pmf
    (pn_kit.py:439-450); the 16-bit conversion and the coder are torchac's (absent: unpinned) and are not stored."""
    pn = ref_loader.load("pn_kit")
    out = {}
    rng = np.random.default_rng(71)
    for name, (n, L, peaky) in {"flat": (512, 7, 0.5), "peaky": (2048, 7, 8.0), "wide": (300, 16, 3.0)}.items():
        pmf = torch.softmax(torch.from_numpy(rng.normal(size=(n, L)).astype(np.float32) * peaky), -1)
        cdf = pn.pmf_to_cdf(pmf)
        sym = torch.from_numpy(rng.integers(0, L, size=(n,)))
        out[f"{name}_pmf"], out[f"{name}_cdf"] = pmf.numpy(), cdf.numpy()
        out[f"{name}_sym"], out[f"{name}_bits"] = sym.numpy(), np.float64(pn.estimate_bits_from_pmf(pmf, sym).item())
        L_ = pmf.shape[-1]
        want = (cdf.mul(2 ** 16 - L_).round().to(torch.int16) + torch.arange(L_ + 1, dtype=torch.int16)).numpy().view(np.uint16)
        assert np.array_equal(orc.pmf_to_cdf_u16(pmf.numpy()), want), name
    np.savez_compressed(os.path.join(HERE, "ref_entropy.npz"), **out)
    print("ref_entropy.npz:", {k: v.shape for k, v in out.items() if k.endswith("_cdf")})


if __name__ == "__main__":
    ref_fps_gather()
    p3d_ops()
    pppf_modules()
    pppe_modules()
    ae_modules()
    ref_octree()
    ref_eval()
    ref_entropy()
