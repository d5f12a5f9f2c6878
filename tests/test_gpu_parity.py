"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the committed golden vectors.
Bit-exact for indices and per-point squared distances; Chamfer scalar within 1e-5 relative (fp32 sum order)."""
import os

import numpy as np
import pytest
import torch

from tools import synth

pytestmark = pytest.mark.gpu

CHAMFER_RTOL = 1e-5  # north_star: "Chamfer within 1e-5 relative"


@pytest.fixture(scope="module")
def pcc():
    import __graft_entry__  # noqa: F401  (sets sys.path)
    import pcc_b200
    assert torch.cuda.is_available()
    pcc_b200._lib.load()
    return pcc_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="module")
def g_ref(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_fps_gather.npz"))


@pytest.fixture(scope="module")
def g_p3d(golden_dir):
    return np.load(os.path.join(golden_dir, "p3d_ops.npz"))


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ---- FPS ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["modelnet", "grid", "perm", "tiny"])
def test_fps_gather_reference_golden(pcc, g_ref, name):
    xyz = cu(g_ref[f"{name}_xyz"])
    idx = pcc.ops.fps(xyz, int(g_ref[f"{name}_npoint"]), cu(g_ref[f"{name}_start"]), 1e10)
    assert np.array_equal(idx.cpu().numpy(), g_ref[f"{name}_idx"])
    assert np.array_equal(pcc.index_points(xyz, idx).cpu().numpy(), g_ref[f"{name}_gather"])
    assert np.array_equal(pcc.index_points(xyz, cu(g_ref[f"{name}_idx3"])).cpu().numpy(), g_ref[f"{name}_gather3"])


def test_fps_reference_rng_protocol(pcc, g_ref):
    """pn_kit.farthest_point_sample_batch draws its start index from the CPU global RNG (pn_kit.py:321)."""
    torch.manual_seed(11)
    idx = pcc.farthest_point_sample_batch(cu(g_ref["modelnet_xyz"]), 64)
    assert np.array_equal(idx.cpu().numpy(), g_ref["modelnet_idx"])


@pytest.mark.parametrize("B,N,S", [(3, 100, 100), (2, 128, 40), (2, 129, 64), (2, 300, 300), (2, 777, 99),
                                   (2, 2048, 512), (1, 4097, 65), (2, 8192, 64), (32, 8192, 64)])
def test_fps_sizes_vs_oracle(pcc, orc, B, N, S):
    xyz = synth.uniform_cube(B, N, seed=N + S) if N % 2 else synth.grid_quantised(B, N, depth=3, seed=N)
    start = np.random.default_rng(N).integers(0, N, B)
    got = pcc.ops.fps(cu(xyz), S, cu(start), 1e10).cpu().numpy()
    assert np.array_equal(got, orc.fps(xyz, S, start, 1e10, threads=8))


def test_fps_multi_cta_cloud(pcc, orc):
    """N > 8192: the cooperative multi-CTA kernel (the 1M-point scene path at a size the oracle finishes fast)."""
    # 3, 2 and 13 CTAs per cloud (slots that carry the winner's coordinates) and 135 (> 128: index-only slots)
    for B, N, S in [(1, 20000, 200), (2, 9000, 64), (1, 100_000, 64), (1, 1_100_000, 24)]:
        xyz = synth.scene_like(N, seed=N)[0][None].repeat(B, 0) if N > 50000 else synth.uniform_cube(B, N, seed=N)
        start = np.arange(B) * 17
        got = pcc.ops.fps(cu(xyz), S, cu(start), 1e10).cpu().numpy()
        assert np.array_equal(got, orc.fps(xyz, S, start, 1e10, threads=8))


def test_fps_multi_cta_iteration_tag_wrap(pcc, orc):
    """The multi-CTA exchange tags its slots with (iteration + 1) & 2047 (fps.cu fps_grid_kernel): more than 2048 iterations
    wrap the tag twice.  N = 20,000 (3 CTAs per cloud), npoint = 4,200, every index against the oracle; a second cloud in the
    batch and grid-quantised points (exact ties) for the tie rule under the wrap."""
    for xyz in (synth.uniform_cube(2, 20000, seed=77), synth.grid_quantised(1, 20000, depth=5, seed=78)):
        start = np.arange(xyz.shape[0]) * 4001 + 7
        got = pcc.ops.fps(cu(xyz), 4200, cu(start), 1e10).cpu().numpy()
        assert np.array_equal(got, orc.fps(xyz, 4200, start, 1e10, threads=8))


def test_fps_multi_cta_18_ctas_ties_and_wrap(pcc, orc):
    """140,000 points = 18 CTAs per cloud, two clouds (7 co-resident clouds per launch would fit: the batch loop), 2,200
    iterations (one tag wrap), grid-quantised points (exact ties everywhere) and uniform ones, every index against the oracle."""
    xyz = np.concatenate((synth.grid_quantised(1, 140_000, depth=6, seed=91), synth.uniform_cube(1, 140_000, seed=92)), axis=0)
    start = np.array([5, 139_999], np.int64)
    got = pcc.ops.fps(cu(xyz), 2200, cu(start), 1e10).cpu().numpy()
    assert np.array_equal(got, orc.fps(xyz, 2200, start, 1e10, threads=8))


class _fps_path:
    """PCC_FPS_PATH ("bucket": fps_bucket.cu whenever it fits; "grid": the co-resident kernel) and PCC_FPS_HEAD (iterations the
    co-resident kernel runs before it hands over to the bucketed form; "0": none) for the duration of a block."""
    def __init__(self, path, head=None):
        self.want = {"PCC_FPS_PATH": path, "PCC_FPS_HEAD": head}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.want}
        for k, v in self.want.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v

    def __exit__(self, *exc):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("head", ["0", None, "37"])
@pytest.mark.parametrize("case", ["scene", "ties", "uniform_pads", "duplicates", "identical", "two_clouds", "tiny"])
def test_fps_bucket_form_vs_oracle(pcc, orc, case, head):
    """The scene-scale form (Morton buckets + exact skipping, fps_bucket.cu) forced on clouds the oracle finishes in seconds:
    every index against the oracle -- exact ties (grid-quantised points), a ragged last bucket, duplicated points (running
    distance 0 early), one point repeated N times, two clouds per launch, the smallest cloud that reaches this path; alone
    (head "0"), after the default head of co-resident iterations (128 at these sizes, when the sampling is >= 512 long) and
    after an odd one (the hand-over of the running distances and of the last picked centre)."""
    if case == "scene":
        xyz, S = synth.scene_like(60_000, seed=5), 3000
    elif case == "ties":
        xyz, S = synth.grid_quantised(1, 70_000, depth=5, seed=7), 2500
    elif case == "uniform_pads":
        xyz, S = synth.uniform_cube(1, 20_001, seed=8), 1500
    elif case == "duplicates":
        base = synth.uniform_cube(1, 3000, seed=9)
        xyz, S = np.tile(base, (1, 5, 1)), 3500          # every point five times: 500 picks at distance 0 at the end
    elif case == "identical":
        xyz, S = np.full((1, 9000, 3), 0.25, np.float32), 40
    elif case == "two_clouds":
        xyz, S = np.concatenate((synth.scene_like(30_000, seed=1), synth.uniform_cube(1, 30_000, seed=2) * 3.0), 0), 1200
    else:
        xyz, S = synth.uniform_cube(1, 8200, seed=3), 300      # just above the single-CTA kernel's 8192 points
    B, N = xyz.shape[:2]
    start = np.array([(7919 * (b + 1)) % N for b in range(B)], np.int64)
    with _fps_path("bucket", head):
        got, got_xyz = pcc.ops.fps(cu(xyz), S, cu(start), 1e10, return_xyz=True)
    got = got.cpu().numpy()
    assert np.array_equal(got, orc.fps(xyz, S, start, 1e10, threads=8))
    assert np.array_equal(got_xyz.cpu().numpy(), np.take_along_axis(xyz, got[:, :, None], 1))


def test_fps_bucket_form_pytorch3d_contract(pcc, orc):
    """sample_farthest_points' contract on the bucketed form: start at index 0, FLT_MAX, -1 padding when K > N."""
    xyz = synth.uniform_cube(2, 9000, seed=4)
    with _fps_path("bucket"):
        idx = pcc.ops.fps(cu(xyz), 9100, None, float(np.finfo(np.float32).max)).cpu().numpy()
    want = orc.fps(xyz, 9000, np.zeros(2, np.int64), float(np.finfo(np.float32).max), threads=8)
    assert np.array_equal(idx[:, :9000], want) and (idx[:, 9000:] == -1).all()


def test_fps_short_batched_sampling_of_big_clouds_takes_the_bucket_form(pcc):
    """Two 620,000-point clouds, 96 centres each: the co-resident kernel fits one such cloud per launch, so the default route is
    the bucketed form (one CTA per cloud, side by side) although the sampling is short -- same indices as the co-resident kernel."""
    xyz = cu(np.concatenate((synth.scene_like(620_000, seed=8), synth.uniform_cube(1, 620_000, seed=9) * 5.0), 0))
    start = cu(np.array([7, 619_999], np.int64))
    a = pcc.ops.fps(xyz, 96, start, 1e10)
    with _fps_path("grid"):
        b = pcc.ops.fps(xyz, 96, start, 1e10)
    assert torch.equal(a, b)


@pytest.mark.parametrize("kind", ["plane", "line", "dups", "offset", "tiny", "lattice", "aniso"])
def test_fps_bucket_form_degenerate_geometry(pcc, orc, kind):
    """Degenerate boxes and arithmetic corners of the skip test (tools/soak_fps.py runs the long version): a plane and a line (zero
    extent along an axis), points repeated seven times, coordinates at 1e6 (fp32 spacing 0.06), an extent of 1e-21 (every square
    underflows: nothing may be skipped), a lattice (ties everywhere), a 10^4 : 1 anisotropic cloud -- against the oracle, alone and
    after a three-iteration head."""
    from tools import soak_fps
    xyz = soak_fps.cloud(kind, 20_000, np.random.default_rng(17))[None]
    start = np.array([4242], np.int64)
    want = orc.fps(xyz, 500, start, 1e10, threads=8)
    for head in ("0", "3"):
        with _fps_path("bucket", head):
            assert np.array_equal(pcc.ops.fps(cu(xyz), 500, cu(start), 1e10).cpu().numpy(), want), head


def test_fps_bucket_form_equals_grid_form_at_scene_scale(pcc):
    """1,000,000 points -> 7812 centres: the default route (244 co-resident iterations, then the bucketed form) against the
    co-resident multi-CTA kernel alone on every iteration (the oracle prefix is test_scene_scale_cfg5's)."""
    xyz = cu(synth.scene_like(1_000_000, seed=3))
    start = cu(np.array([12345], np.int64))
    a = pcc.ops.fps(xyz, 7812, start, 1e10)
    with _fps_path("grid"):
        b = pcc.ops.fps(xyz, 7812, start, 1e10)
    assert torch.equal(a, b)


@pytest.mark.parametrize("name", ["sfp", "sfp_pad", "sfp_ties"])
def test_sample_farthest_points_golden(pcc, g_p3d, name):
    pts, idx = pcc.sample_farthest_points(cu(g_p3d[f"{name}_x"]), K=int(g_p3d[f"{name}_K"]))
    assert np.array_equal(idx.cpu().numpy(), g_p3d[f"{name}_idx"])
    assert np.array_equal(pts.cpu().numpy(), g_p3d[f"{name}_pts"])


# ---- kNN ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["knn_patch", "knn_dupq", "knn_ties", "knn_self", "knn_kgtn"])
def test_knn_golden(pcc, g_p3d, name):
    r = pcc.knn_points(cu(g_p3d[f"{name}_q"]), cu(g_p3d[f"{name}_p"]), K=int(g_p3d[f"{name}_K"]), return_nn=True)
    dists, idx, nn = r  # the reference unpacks the namedtuple as a 3-tuple (train.py:185)
    assert np.array_equal(r.idx.cpu().numpy(), g_p3d[f"{name}_i"])
    assert np.array_equal(dists.cpu().numpy(), g_p3d[f"{name}_d"])
    assert np.array_equal(nn.cpu().numpy(), g_p3d[f"{name}_nn"])


@pytest.mark.parametrize("B,P1,P2,K", [(1, 1, 5, 3), (2, 64, 8192, 256), (3, 33, 1500, 100), (2, 7, 5000, 1024),
                                       (1, 1, 8192, 1024), (2, 512, 2048, 16), (2, 512, 2048, 32), (2, 100, 40, 64),
                                       (1, 9, 3000, 1), (4, 130, 1025, 17)])
def test_knn_warp_kernel_vs_oracle(pcc, orc, B, P1, P2, K):
    p = synth.grid_quantised(B, P2, depth=4, seed=P2) if (P1 + K) % 2 else synth.uniform_cube(B, P2, seed=P2)
    q = synth.uniform_cube(B, P1, seed=P1 + 1)
    d, i, nn = pcc.ops.knn(cu(q), cu(p), K, return_nn=True, centre_sub=True, nn_scale=2.0)
    od, oi, onn = orc.knn_points(q, p, K, True, threads=8)
    assert np.array_equal(i.cpu().numpy(), oi) and np.array_equal(d.cpu().numpy(), od)
    want = (onn - q[:, :, None, :]) * np.float32(2.0)
    if P2 >= K:
        assert np.array_equal(nn.cpu().numpy(), want)


@pytest.mark.parametrize("kind,P1,P2,K", [("uniform", 40, 20000, 256), ("grid", 17, 9001, 64), ("grid", 9, 30000, 256),
                                          ("identical", 3, 10000, 100), ("uniform", 5, 12000, 1024), ("two_values", 6, 8300, 200),
                                          ("uniform", 33, 8193, 16)])
def test_knn_scene_kernel_vs_oracle(pcc, orc, kind, P1, P2, K):
    """P2 > 8192: the warp-per-query kernel of the scene path -- 128-candidate steps with the one-vote reject, the ragged last
    tile, the bisection prune of a full candidate buffer, its fall-back to the exact sort when ties leave no room (grid /
    identical / two_values) and the sort-only path of K > 256."""
    rng = np.random.default_rng(P2 + K)
    if kind == "identical":
        p = np.full((1, P2, 3), 0.25, np.float32)
    elif kind == "two_values":
        p = np.where(rng.random((1, P2, 1)) < 0.5, 0.25, 0.75).astype(np.float32).repeat(3, axis=2)
    elif kind == "grid":
        p = synth.grid_quantised(1, P2, depth=3, seed=P2)
    else:
        p = synth.uniform_cube(1, P2, seed=P2)
    q = synth.uniform_cube(1, P1, seed=P1 + 1)
    q[0, 0] = p[0, 7]                                                   # a query sitting on a candidate
    d, i, nn = pcc.ops.knn(cu(q), cu(p), K, return_nn=True)
    od, oi, onn = orc.knn_points(q, p, K, True, threads=8)
    assert np.array_equal(i.cpu().numpy(), oi) and np.array_equal(d.cpu().numpy(), od)
    assert np.array_equal(nn.cpu().numpy(), onn)


@pytest.mark.parametrize("kind,P1,P2,K", [("scene", 300, 200_000, 256), ("uniform", 64, 70_000, 16), ("grid", 40, 100_000, 256),
                                          ("identical", 5, 66_000, 100), ("outside", 33, 80_000, 32), ("scene", 9, 300_000, 1024),
                                          ("clumps", 50, 120_000, 1), ("small", 20, 3000, 64)])
def test_knn_grid_form_vs_brute_force_and_oracle(pcc, orc, kind, P1, P2, K):
    """pcc_knn_grid_f32 (scene scale: cell, then shell after shell of a 32^3 grid until the K-th distance is inside the scanned
    block) against the brute-force warp kernel on the same inputs -- distances, indices and gathered neighbours must be identical
    -- and against the CPU oracle on a few queries: surface-like scenes, exact ties (grid-quantised, all-identical points),
    queries outside the cloud's bounding box, far clumps (whole shells empty), K from 1 to 1024, and a cloud smaller than the
    grid's cell count."""
    rng = np.random.default_rng(P2 + K)
    if kind == "scene":
        p = synth.scene_like(P2, seed=P2)
    elif kind == "grid":
        p = synth.grid_quantised(1, P2, depth=5, seed=P2)
    elif kind == "identical":
        p = np.full((1, P2, 3), 0.25, np.float32)
    elif kind == "clumps":
        c = rng.uniform(0, 10, (6, 3)).astype(np.float32)
        p = (c[rng.integers(0, 6, P2)] + rng.normal(0, 0.01, (P2, 3)).astype(np.float32))[None]
    else:
        p = synth.uniform_cube(1, P2, seed=P2)
    q = p[:, rng.integers(0, P2, P1)].copy()
    if kind == "outside":
        q = (q * 3.0 - 1.0).astype(np.float32)
    if kind == "clumps":
        q[0, ::2] += np.float32(3.0)
    q[0, 0] = p[0, 7]
    d, i, nn = pcc.ops.knn(cu(q), cu(p), K, return_nn=True, centre_sub=True, nn_scale=2.0, grid=True)
    bd, bi, bnn = pcc.ops.knn(cu(q), cu(p), K, return_nn=True, centre_sub=True, nn_scale=2.0, grid=False)
    assert torch.equal(i, bi) and torch.equal(d, bd) and torch.equal(nn, bnn)
    n = min(P1, 6)
    od, oi, _ = orc.knn_points(q[:, :n], p, K, False, threads=8)
    assert np.array_equal(i[:, :n].cpu().numpy(), oi) and np.array_equal(d[:, :n].cpu().numpy(), od)


def test_scene_patches_single_process_equals_direct_ops(pcc, orc):
    """dist.scene_patches (cfg5: FPS on rank 0 + broadcast, kNN queries split over ranks + all-gather) in a single process is
    the plain FPS + kNN; its FPS prefix is checked against the oracle, the returned local patches against the gathered table."""
    from pcc_b200 import dist as pdist
    xyz = cu(synth.scene_like(150_000, seed=31))
    start = torch.tensor([4321], dtype=torch.int64)
    fps_idx, knn_idx, (b, e, nn) = pdist.scene_patches(xyz, 600, 64, start, return_local_nn=True)
    assert (b, e) == (0, 600) and fps_idx.shape == (1, 600) and knn_idx.shape == (1, 600, 64) and nn.shape == (1, 600, 64, 3)
    assert np.array_equal(fps_idx[0, :120].cpu().numpy(), orc.fps(xyz.cpu().numpy(), 120, start.numpy(), 1e10, threads=8)[0])
    cen = pcc.ops.gather(xyz, fps_idx)
    d, i, want_nn = pcc.ops.knn(cen, xyz, 64, return_nn=True, centre_sub=True)
    assert torch.equal(i, knn_idx) and torch.equal(want_nn, nn)
    again = pdist.scene_patches(xyz, 600, 64, fps_idx=fps_idx)          # the kNN leg alone (bench cfg5 times it this way)
    assert torch.equal(again[0], fps_idx) and torch.equal(again[1], knn_idx)


@pytest.mark.parametrize("kind,P2,K", [("identical", 8192, 256), ("two_values", 8192, 300), ("grid", 8192, 512),
                                       ("grid", 5000, 64), ("uniform", 4097, 512), ("half_dup", 6000, 256)])
def test_knn_block_kernel_selection_paths_vs_oracle(pcc, orc, kind, P2, K):
    """CTA-per-query kernel (4096 < P2 <= 8192): the sort-based selection (per-thread minima bound + admitted keys) and its
    fall-back to the bisection when ties let more than 512 candidates through -- identical results either way."""
    rng = np.random.default_rng(P2 + K)
    if kind == "identical":
        p = np.full((2, P2, 3), 0.25, np.float32)                      # every distance ties: admitted = P2 -> bisection path
    elif kind == "two_values":
        p = np.where(rng.random((2, P2, 1)) < 0.5, 0.25, 0.75).astype(np.float32).repeat(3, axis=2)
    elif kind == "grid":
        p = synth.grid_quantised(2, P2, depth=3, seed=P2)              # heavy ties at the K-th distance
    elif kind == "half_dup":
        p = synth.uniform_cube(2, P2, seed=P2)
        p[:, P2 // 2:] = p[:, :P2 - P2 // 2]                            # every point twice: tie pairs ordered by index
    else:
        p = synth.uniform_cube(2, P2, seed=P2)
    q = synth.uniform_cube(2, 5, seed=K)
    q[:, 0] = p[:, 17]                                                  # a query that coincides with a candidate
    d, i, _ = pcc.ops.knn(cu(q), cu(p), K)
    od, oi, _ = orc.knn_points(q, p, K, False, threads=8)
    assert np.array_equal(i.cpu().numpy(), oi) and np.array_equal(d.cpu().numpy(), od)


@pytest.mark.parametrize("kind", ["uniform", "grid", "identical", "offset", "tiny", "line", "two_clusters", "qp_differ", "near_dup"])
def test_knn_in_patch_adversarial_vs_oracle(pcc, orc, kind):
    """256 x 256, K = 16 (pn_kit.py:190 as AE.py:16 calls it) on inputs that stress the a-priori threshold and the tie rules:
    exact ties, all-identical points, large offsets, tiny scales, collinear points, two far clusters, queries != candidates."""
    BS = 160   # 160 * 256 queries >= the dispatch threshold of the batched kernels
    rng = np.random.default_rng(len(kind))
    x = synth.uniform_cube(BS, 256, seed=7).astype(np.float32) - np.float32(0.5)
    if kind == "grid":
        x = synth.grid_quantised(BS, 256, depth=3, seed=8) - np.float32(0.5)           # exact ties everywhere
    elif kind == "identical":
        x[:] = x[:, :1]                                                               # all distances 0: order by index
    elif kind == "offset":
        x = x + np.float32(100.0)                                                     # |x|^2 ~ 3e4: the margin admits almost everything
    elif kind == "tiny":
        x = x * np.float32(1e-4)
    elif kind == "line":
        x[:, :, 1:] = 0                                                               # collinear points
    elif kind == "two_clusters":
        x[:, ::2] *= np.float32(1e-3)
        x[:, 1::2] = x[:, 1::2] * np.float32(1e-3) + np.float32(1.0)
    elif kind == "near_dup":
        x[:, 1::4] = x[:, 0::4] + rng.normal(0, 1e-6, x[:, 0::4].shape).astype(np.float32)   # d2 ~ 1e-12 beside d2 ~ 1e-2: the
        #                                                     packed 27-bit key range of the filter kernel does not hold them
    q = x
    if kind == "qp_differ":
        q = np.ascontiguousarray(x[:, ::-1]) * np.float32(0.9)
    d, i, nn = pcc.ops.knn(cu(q), cu(x), 16, return_nn=True, centre_sub=True)
    od, oi, onn = orc.knn_points(q, x, 16, True, threads=8)
    assert np.array_equal(i.cpu().numpy(), oi) and np.array_equal(d.cpu().numpy(), od)
    assert np.array_equal(nn.cpu().numpy(), onn - q[:, :, None, :])


@pytest.mark.parametrize("BS,P,K", [(200, 256, 16), (300, 128, 8), (150, 256, 32), (140, 250, 20), (600, 64, 3)])
def test_knn_thread_kernel_in_patch_vs_oracle(pcc, orc, BS, P, K):
    """Many small self-searches (the pn_kit.SetAbstraction K=16 case) -> thread-per-query kernel."""
    x = synth.grid_quantised(BS, P, depth=3, seed=K) if K == 16 else synth.uniform_cube(BS, P, seed=K)
    d, i, nn = pcc.ops.knn(cu(x), cu(x), K, return_nn=True, centre_sub=True)
    od, oi, onn = orc.knn_points(x, x, K, True, threads=8)
    assert np.array_equal(i.cpu().numpy(), oi) and np.array_equal(d.cpu().numpy(), od)
    assert np.array_equal(nn.cpu().numpy(), onn - x[:, :, None, :])


@pytest.mark.parametrize("BS,P,K,kind", [(64, 256, 16, "uniform"), (2048, 256, 16, "patches"), (40, 128, 8, "grid"), (7, 200, 16, "identical"),
                                         (3, 16, 16, "uniform"), (33, 256, 8, "near_dup")])
def test_knn_patch_u8_vs_oracle(pcc, orc, BS, P, K, kind):
    """pcc_knn_patch_u8 (the byte index table the indexed SetAbstraction chain reads): equal to the oracle's knn_points indices
    on small / tie-heavy / degenerate patches and on the bench's real patches (32 clouds x 64 patches)."""
    if kind == "patches":
        xyz = cu(synth.modelnet_like(32, 8192, seed=4))
        cen = pcc.index_points(xyz, pcc.ops.fps(xyz, 64, None, 1e10))
        x = pcc.ops.knn(cen, xyz, 256, return_nn=True, centre_sub=True, nn_scale=2.0)[2].reshape(BS, P, 3).cpu().numpy()
    elif kind == "grid":
        x = synth.grid_quantised(BS, P, depth=3, seed=9) - np.float32(0.5)
    else:
        x = synth.uniform_cube(BS, P, seed=10) - np.float32(0.5)
        if kind == "identical":
            x[:] = x[:, :1]
        if kind == "near_dup":
            x[:, 1::4] = x[:, 0::4] + np.random.default_rng(3).normal(0, 1e-6, x[:, 0::4].shape).astype(np.float32)
    i8 = pcc.ops.knn_patch_u8(cu(x), K)
    assert i8.dtype == torch.uint8 and i8.shape == (BS, P, K)
    _, oi, _ = orc.knn_points(x, x, K, False, threads=8)
    assert np.array_equal(i8.cpu().numpy().astype(np.int64), oi)
    with pytest.raises(ValueError):
        pcc.ops.knn_patch_u8(cu(x), 12)                      # K must be 8 or 16
    with pytest.raises(ValueError):
        pcc.ops.knn_patch_u8(cu(synth.uniform_cube(2, 300, seed=1)), 16)   # byte indices: P <= 256


def test_knn_patching_full_size_properties(pcc):
    """BASELINE size (32 clouds x 64 centres x 8192 points, K=256): size-independent properties."""
    xyz = cu(synth.modelnet_like(32, 8192, seed=3))
    centres = pcc.index_points(xyz, pcc.ops.fps(xyz, 64, None, 1e10))
    d, i, nn = pcc.ops.knn(centres, xyz, 256, return_nn=True)
    assert bool((d[:, :, 1:] >= d[:, :, :-1]).all())  # ascending
    same = d[:, :, 1:] == d[:, :, :-1]
    assert bool((i[:, :, 1:][same] > i[:, :, :-1][same]).all())  # ties ordered by index
    assert bool((d[:, :, 0] == 0).all()) and bool((nn[:, :, 0] == centres).all())  # a centre is its own 1st NN
    assert torch.equal(nn, pcc.index_points(xyz, i))  # gathered neighbours == gather(idx)
    d2 = ((nn - centres[:, :, None, :]) ** 2)
    assert torch.equal(d, (d2[..., 0] + d2[..., 1]) + d2[..., 2])  # distances re-derivable, un-fused
    d1, i1 = pcc.ops.nn1(centres, xyz)
    assert torch.equal(d1, d[:, :, 0]) and torch.equal(i1, i[:, :, 0])  # K=1 search agrees with the K=256 head


# ---- ball query / gather ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["ball_sa1", "ball_edge", "ball_few"])
def test_ball_query_golden(pcc, g_p3d, name):
    r = pcc.ball_query(cu(g_p3d[f"{name}_q"]), cu(g_p3d[f"{name}_p"]), K=int(g_p3d[f"{name}_K"]),
                       radius=float(g_p3d[f"{name}_r"]))
    assert np.array_equal(r.idx.cpu().numpy(), g_p3d[f"{name}_i"])
    assert np.array_equal(r.dists.cpu().numpy(), g_p3d[f"{name}_d"])
    want = g_p3d[f"{name}_p"][np.arange(r.idx.shape[0])[:, None, None], np.maximum(g_p3d[f"{name}_i"], 0)]
    want[g_p3d[f"{name}_i"] < 0] = 0
    assert np.array_equal(r.knn.cpu().numpy(), want)


@pytest.mark.parametrize("B,P1,P2,K,r", [(4, 512, 2048, 32, 0.2), (4, 128, 512, 64, 0.4), (4, 32, 128, 128, 0.8),
                                         (2, 5, 3000, 7, 0.05), (1, 1, 1, 4, 1.0)])
def test_ball_query_pointnetpp_shapes_vs_oracle(pcc, orc, B, P1, P2, K, r):
    p = synth.shapenet_like(B, P2, seed=P2)
    q = p[:, :P1] if P1 <= P2 else synth.uniform_cube(B, P1)
    idx = pcc.PointnetPPOps.ball_query(r, K, cu(p), cu(q))  # (radius, nsample, xyz, new_xyz) as the reference
    od, oi = orc.ball_query(q, p, K, r, threads=8)
    assert np.array_equal(idx.idx.cpu().numpy(), oi) and np.array_equal(idx.dists.cpu().numpy(), od)
    feats = synth.uniform_cube(B, P2 * 16, seed=1).reshape(B, P2, 48)
    grouped = pcc.PointnetPPOps.group_points(cu(feats), idx)  # -1 pads clamp to point 0 (pointnet_sa_module.py:27)
    assert np.array_equal(grouped.cpu().numpy(), orc.gather(feats, np.maximum(oi, 0)))


def test_gather_backward_is_scatter_add(pcc):
    feat = torch.rand(2, 50, 8, device="cuda", requires_grad=True)
    idx = torch.randint(0, 50, (2, 30, 4), device="cuda")
    out = pcc.knn_gather(feat, idx)
    w = torch.rand_like(out)
    (out * w).sum().backward()
    ref = torch.zeros(2, 50, 8, device="cuda")
    ref.scatter_add_(1, idx.reshape(2, -1, 1).expand(-1, -1, 8), w.reshape(2, -1, 8))
    assert torch.allclose(feat.grad, ref, atol=1e-6)


# ---- Chamfer / D1 -----------------------------------------------------------------------------------------------
def test_chamfer_golden(pcc, g_p3d):
    x, y = cu(g_p3d["cham_x"]), cu(g_p3d["cham_y"])
    r = pcc.ops.chamfer_forward(x, y)
    assert np.array_equal(r["dx"].cpu().numpy(), g_p3d["cham_dx"]) and np.array_equal(r["ix"].cpu().numpy(), g_p3d["cham_ix"])
    assert np.array_equal(r["dy"].cpu().numpy(), g_p3d["cham_dy"]) and np.array_equal(r["iy"].cpu().numpy(), g_p3d["cham_iy"])
    want = float(g_p3d["cham_loss"])
    assert abs(r["loss"].item() - want) <= CHAMFER_RTOL * abs(want)
    assert np.allclose(r["per_cloud"].cpu().numpy(), g_p3d["cham_pc"], rtol=CHAMFER_RTOL, atol=0)
    xg, yg = x.clone().requires_grad_(), y.clone().requires_grad_()
    loss, normals = pcc.chamfer_distance(xg, yg)
    assert normals is None
    loss.backward()
    assert np.allclose(xg.grad.cpu().numpy(), g_p3d["cham_gx"], rtol=1e-5, atol=1e-9)
    assert np.allclose(yg.grad.cpu().numpy(), g_p3d["cham_gy"], rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("B,P1,P2", [(1, 8192, 8192), (3, 1000, 777), (2, 16384, 8192), (1, 1, 1), (5, 9, 2050)])
def test_chamfer_vs_oracle(pcc, orc, B, P1, P2):
    x = synth.modelnet_like(B, max(P1, P2), seed=P1)
    y = synth.decompressed_like(x, seed=P2)[:, :P2]
    x = x[:, :P1]
    if P1 == 1000:  # tie-heavy
        x, y = synth.grid_quantised(B, P1, depth=3, seed=1), synth.grid_quantised(B, P2, depth=3, seed=2)
    r = pcc.ops.chamfer_forward(cu(x), cu(y))
    loss, pc, dx, ix, dy, iy = orc.chamfer(x, y, threads=8)
    assert np.array_equal(r["dx"].cpu().numpy(), dx) and np.array_equal(r["dy"].cpu().numpy(), dy)
    assert np.array_equal(r["ix"].cpu().numpy(), ix) and np.array_equal(r["iy"].cpu().numpy(), iy)
    assert abs(r["loss"].item() - loss) <= CHAMFER_RTOL * abs(loss) + 1e-30


def _adversarial_pair(kind, P1, P2, seed):
    rng = np.random.default_rng(seed)
    if kind == "ties":          # octree-grid points: exact ties and duplicates everywhere
        return synth.grid_quantised(1, P1, depth=4, seed=seed), synth.grid_quantised(1, P2, depth=4, seed=seed + 1)
    if kind == "identical":     # zero-extent clouds
        x = np.full((1, P1, 3), 0.25, np.float32)
        y = np.full((1, P2, 3), 0.25, np.float32)
        y[0, 7] = (0.5, 0.25, 0.25)
        return x, y
    if kind == "outliers":      # a tight cluster plus a few far points: almost every point in one cell
        x = (rng.normal(0, 1e-3, (1, P1, 3)) + 0.5).astype(np.float32)
        y = (rng.normal(0, 1e-3, (1, P2, 3)) + 0.5).astype(np.float32)
        x[0, :5] = rng.uniform(-50, 50, (5, 3))
        y[0, :3] = rng.uniform(-50, 50, (3, 3))
        return x, y
    if kind == "offset":        # large common offset: coordinates ~1000, extent ~1 (face rounding vs the margin)
        x = synth.modelnet_like(1, P1, seed=seed) + np.float32(1000.0)
        y = synth.decompressed_like(x, seed=seed + 1)[:, :P2]
        return x.astype(np.float32), y.astype(np.float32)
    if kind == "disjoint":      # the two clouds do not overlap at all: every query leaves the other grid
        x = synth.modelnet_like(1, P1, seed=seed)
        y = synth.modelnet_like(1, P2, seed=seed + 1) + np.array([3.0, -2.0, 0.5], np.float32)
        return x, y.astype(np.float32)
    if kind == "line":          # points on a segment (one occupied row of cells) against a surface
        t = rng.uniform(0, 1, (1, P1, 1)).astype(np.float32)
        x = np.concatenate((t, 0.3 * t + 0.1, 0.5 - 0.2 * t), axis=2).astype(np.float32)
        return x, synth.modelnet_like(1, P2, seed=seed)
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["ties", "identical", "outliers", "offset", "disjoint", "line"])
@pytest.mark.parametrize("P1,P2", [(4096, 4096), (1024, 2500), (2049, 1025)])
def test_chamfer_grid_path_adversarial_vs_oracle(pcc, orc, kind, P1, P2):
    """The grid-pruned search (clouds >= 1024 points) must return exactly what the exhaustive search returns."""
    x, y = _adversarial_pair(kind, P1, P2, seed=P1 + len(kind))
    r = pcc.ops.chamfer_forward(cu(x), cu(y))
    loss, pc, dx, ix, dy, iy = orc.chamfer(x, y, threads=8)
    assert np.array_equal(r["dx"].cpu().numpy(), dx) and np.array_equal(r["dy"].cpu().numpy(), dy)
    assert np.array_equal(r["ix"].cpu().numpy(), ix) and np.array_equal(r["iy"].cpu().numpy(), iy)
    assert abs(r["loss"].item() - loss) <= CHAMFER_RTOL * abs(loss) + 1e-30
    # the distance-only form (no index wanted: packed-fp32 loop over paired candidates) returns the same minima, bit for bit
    n = pcc.ops.chamfer_forward(cu(x), cu(y), want_idx=False)
    assert np.array_equal(n["dx"].cpu().numpy(), dx) and np.array_equal(n["dy"].cpu().numpy(), dy)
    assert torch.equal(n["per_cloud"], r["per_cloud"]) and torch.equal(n["loss"], r["loss"])


def test_chamfer_full_batch_properties(pcc):
    """BASELINE size (32 x 8192 x 8192): symmetry, zero self-distance, agreement with the 1-NN entry point."""
    x = cu(synth.modelnet_like(32, 8192, seed=7))
    y = cu(synth.decompressed_like(x.cpu().numpy(), seed=8))
    a = pcc.ops.chamfer_forward(x, y)
    b = pcc.ops.chamfer_forward(y, x)
    assert torch.equal(a["dx"], b["dy"]) and torch.equal(a["ix"], b["iy"]) and torch.equal(a["per_cloud"], b["per_cloud"])
    z = pcc.ops.chamfer_forward(x, x)
    assert float(z["loss"]) == 0.0 and torch.equal(z["ix"], torch.arange(8192, device="cuda").expand(32, -1))
    d, i = pcc.ops.nn1(x, y)
    assert torch.equal(d, a["dx"]) and torch.equal(i, a["ix"])
    assert abs(a["per_cloud"].double().mean().item() - a["loss"].item()) <= 1e-6 * a["loss"].item()
    n = pcc.ops.chamfer_forward(x, y, want_idx=False)     # eval.py's form: distances only
    assert torch.equal(n["dx"], a["dx"]) and torch.equal(n["dy"], a["dy"]) and torch.equal(n["per_cloud"], a["per_cloud"])
    assert n["ix"] is None and n["iy"] is None


def test_d1_psnr_inner_step(pcc, orc, g_p3d):
    """eval.py:68-92: PSNR from the one-directional 1-NN (recon -> original); fp32 device NN vs the float64 oracle."""
    orig, recon = g_p3d["cham_x"][0], g_p3d["cham_y"][0]
    d, _ = pcc.ops.nn1(cu(recon[None]), cu(orig[None]))
    mse = d.double().mean().item()
    diag2 = float(((orig.max(0) - orig.min(0)).astype(np.float64) ** 2).sum())
    psnr = 10 * np.log10(diag2 / mse)
    assert abs(psnr - float(g_p3d["d1_psnr"])) < 1e-4  # dB


# ---- fused glue kernels of the batched driver -------------------------------------------------------------------------
def test_normalize_matches_pn_kit_formula(pcc):
    """pcc_normalize_f32 vs the reference op order of pn_kit.normalize (pn_kit.py:47-60), applied per cloud."""
    x = cu(synth.modelnet_like(3, 8192, seed=41) * np.float32(3.7) - np.float32(1.2))
    out, center, longest, bbox = pcc.ops.normalize(x)
    for b in range(3):
        pc = x[b:b + 1]
        xs, ys, zs = pc[0, :, 0], pc[0, :, 1], pc[0, :, 2]
        c = torch.stack(((xs.max() + xs.min()) / 2, (ys.max() + ys.min()) / 2, (zs.max() + zs.min()) / 2))
        l = torch.max(torch.stack((xs.max() - xs.min(), ys.max() - ys.min(), zs.max() - zs.min())))
        ref = (pc - c) * (1 - 0.01) / l + 0.5
        assert torch.equal(center[b], c) and torch.equal(longest[b], l)
        assert torch.equal(out[b:b + 1], ref)
        assert torch.equal(bbox[b], torch.cat((pc[0].amin(0), pc[0].amax(0))))


def test_fps_fused_centres_and_quantisation(pcc, orc):
    xyz = synth.modelnet_like(2, 8192, seed=43)
    start = np.array([11, 4000])
    idx, cen = pcc.ops.fps(cu(xyz), 64, cu(start), 1e10, return_xyz=True)
    assert np.array_equal(idx.cpu().numpy(), orc.fps(xyz, 64, start, 1e10))
    assert np.array_equal(cen.cpu().numpy(), orc.gather(xyz, idx.cpu().numpy()))
    _, q = pcc.ops.fps(cu(xyz), 64, cu(start), 1e10, return_xyz=True, quant_cube=1.0 / 64)
    cube = np.float32(1.0 / 64)
    want = (orc.gather(xyz, idx.cpu().numpy()) // cube * cube) + cube / 2   # octree_np.getDecodeFromPc's rule
    assert np.array_equal(q.cpu().numpy(), want.astype(np.float32))


def test_assemble_matches_reference_ops(pcc):
    B, S, k = 2, 64, 128
    patches = torch.rand(B * S, k, 3, device="cuda") - 0.5
    centres = torch.rand(B, S, 3, device="cuda")
    center = torch.rand(B, 3, device="cuda")
    longest = torch.rand(B, device="cuda") + 0.5
    got = pcc.ops.assemble(patches, centres, 2.0, center, longest)
    pc = (patches / 2.0).view(B, S, k, 3) + centres.view(B, S, 1, 3)   # decompress.py:105-110
    pc = pc.reshape(B, -1, 3)
    ref = (pc - 0.5) * longest[:, None, None] / (1 - 0.01) + center[:, None, :]  # pn_kit.denormalize
    assert torch.allclose(got, ref, rtol=0, atol=1e-6)
    assert torch.equal(pcc.ops.assemble(patches, centres, 2.0), pc)


# ---- octree centre coding: bit-exact against the REFERENCE's own coder and against the oracle ----------
@pytest.fixture(scope="module")
def g_oct(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_octree.npz"))


@pytest.mark.parametrize("name", ["k256", "k128", "k1024", "s300", "dup", "grid", "one", "edge"])
def test_octree_encode_reference_golden(pcc, g_oct, name):
    c, N, min_bpp = g_oct[f"{name}_c"], int(g_oct[f"{name}_N"]), float(g_oct[f"{name}_min_bpp"])
    r = pcc.ops.octree_encode(cu(c), N, min_bpp, 0, want_bytes=True, want_quant=True, want_rec_ref=True, want_stream_xyz=True)
    nbits, depth = r["nbits"].cpu().numpy(), r["depth"].cpu().numpy()
    assert np.array_equal(nbits, g_oct[f"{name}_nbits"]) and np.array_equal(depth, g_oct[f"{name}_depth"])
    bits, by = r["bits"].cpu().numpy(), r["bytes"].cpu().numpy()
    B, S, _ = c.shape
    for b in range(B):
        n = int(nbits[b])
        assert np.array_equal(bits[b, :n], g_oct[f"{name}_bits"][b, :n]) and not bits[b, n:].any()
        assert np.array_equal(by[b, :(n + 7) // 8], g_oct[f"{name}_bytes"][b, :(n + 7) // 8])
        # snapped centres: as a set they are the reference's getDecodeFromPc (np.unique rows); the stream-order list is
        # the same set in descending (x, y, z) cell order with the last leaf repeated
        u = g_oct[f"{name}_uniq{b}"]
        q = r["quant"][b].cpu().numpy()
        assert np.array_equal(np.unique(q, axis=0), u)
        sx = r["stream_xyz"][b].cpu().numpy()
        cells = np.floor(u * np.float32(2.0 ** depth[b])).astype(np.int64)
        morton = np.zeros(len(u), np.int64)
        for lvl in range(int(depth[b]) - 1, -1, -1):  # child index 4x + 2y + z per level, coarsest level most significant
            morton = (morton << 3) | (((cells[:, 0] >> lvl) & 1) << 2) | (((cells[:, 1] >> lvl) & 1) << 1) | ((cells[:, 2] >> lvl) & 1)
        assert np.array_equal(sx[:len(u)], u[np.argsort(-morton, kind="stable")]) and np.all(sx[len(u):] == sx[len(u) - 1])
    assert np.array_equal(r["rec_ref"].cpu().numpy(), g_oct[f"{name}_rec"])
    # decoders: mode 0 = the reference's decode_sampled_np, mode 1 = inverse of the encoder
    rec0, _, _ = pcc.ops.octree_decode(r["bits"], r["nbits"], mode=0, cap=64)
    assert np.array_equal(rec0.cpu().numpy(), g_oct[f"{name}_rec"])
    rec1, count, d1 = pcc.ops.octree_decode(r["bits"], r["nbits"], mode=1, cap=S)
    assert np.array_equal(rec1.cpu().numpy(), r["stream_xyz"].cpu().numpy())
    assert np.array_equal(d1.cpu().numpy(), depth)
    assert count.cpu().numpy().tolist() == [len(g_oct[f"{name}_uniq{b}"]) for b in range(B)]


def test_octree_fixed_depth_and_drop_in_names(pcc, g_oct):
    from pcc_b200 import octree_ops
    c = g_oct["k256_c"]
    for d in (1, 2, 5, 9):
        assert np.array_equal(octree_ops.encode(c[0], 1, d), g_oct[f"fixed_d{d}"])
    codes, total = octree_ops.encode_sampled_np(c, scale=1, N=8192, min_bpp=0.25)       # pn_kit.py:380 signature, numpy in
    assert total == int(g_oct["k256_nbits"].sum())
    assert all(np.array_equal(code, g_oct["k256_bits"][b, :len(code)]) for b, code in enumerate(codes))
    assert np.array_equal(octree_ops.decode_sampled_np(codes, scale=1), g_oct["k256_rec"])
    assert np.array_equal(octree_ops.decode(codes[0], 1), g_oct["k256_rec"][0])
    with pytest.raises(NotImplementedError):
        octree_ops.encode_sampled_np(c, scale=2, N=8192, min_bpp=0.25)
    bad = c.copy()
    bad[1, 3, 0] = 1.0
    with pytest.raises(ValueError):
        octree_ops.encode_sampled_np(bad, scale=1, N=8192, min_bpp=0.25)


@pytest.mark.parametrize("B,S,N,min_bpp", [(4, 2, 64, 0.3), (3, 63, 2048, 0.25), (2, 257, 8192, 0.2), (1, 1000, 16384, 0.5),
                                           (1, 2500, 100000, 0.3), (1, 7812, 1000000, 0.05)])
def test_octree_encode_vs_oracle(pcc, orc, B, S, N, min_bpp):
    c = synth.uniform_cube(B, S, seed=S)
    if S == 257:
        c[:, 200:] = c[:, :57]  # duplicated centres
    r = pcc.ops.octree_encode(cu(c), N, min_bpp, 0, want_bytes=True)
    codes, total, depths = orc.encode_sampled_np(c, 1, N, min_bpp)
    assert r["depth"].cpu().numpy().tolist() == depths
    bits = r["bits"].cpu().numpy()
    for b, code in enumerate(codes):
        assert int(r["nbits"][b]) == len(code) and np.array_equal(bits[b, :len(code)], code)
        assert np.array_equal(r["bytes"][b, :(len(code) + 7) // 8].cpu().numpy(), orc.bits_to_bytes(code))


# ---- eval.py's remaining metrics: uniformity coefficient and point-to-plane PSNR -------------------------
def test_uniformity_coefficient_and_p2plane(pcc, orc, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_eval.npz"))
    x = synth.modelnet_like(3, 8192, seed=81)
    y = synth.decompressed_like(x, seed=82)
    uc = pcc.ops.uniformity_coefficient(cu(x), cu(y)).cpu().numpy()
    for b in range(3):
        assert abs(uc[b] - orc.calc_uc(x[b], y[b])) <= 1e-9 * uc[b]       # same d2 bit patterns, double reductions
        assert abs(uc[b] - g["uc"][b]) <= 1e-4 * g["uc"][b]              # the reference's own calc_uc (cdist matmul path)
    pp = pcc.ops.p2plane_psnr(cu(y), cu(x)).cpu().numpy()                 # (recon, orig)
    for b in range(3):
        assert abs(pp[b, 1] - g["p2plane"][b, 0]) < 1e-3                  # dB; analytic eigenvectors vs numpy eigh
        assert abs(pp[b, 0] - g["p2plane"][b, 1]) <= 1e-4 * g["p2plane"][b, 1]
    # normals: unit length, orthogonal to the local surface of a plane
    plane = np.random.default_rng(3).random((1, 2000, 3)).astype(np.float32)
    plane[..., 2] = 0.25
    n = pcc.ops.estimate_normals(cu(plane)).cpu().numpy()
    assert np.allclose(np.abs(n[..., 2]), 1.0, atol=1e-5)
    nx = pcc.ops.estimate_normals(cu(x[:1])).cpu().numpy()[0]
    ref = orc.estimate_normals(x[0])
    assert np.median(np.abs(np.abs((nx * ref).sum(-1)) - 1.0)) < 1e-6     # same direction up to sign


def test_empty_and_degenerate_inputs(pcc):
    """Zero-sized batches / query sets return empty tensors (no launch), single points work, bad shapes raise."""
    dev = "cuda"
    p = torch.rand(2, 50, 3, device=dev)
    d, i, nn = pcc.ops.knn(torch.empty(2, 0, 3, device=dev), p, 4, return_nn=True)
    assert d.shape == (2, 0, 4) and i.shape == (2, 0, 4) and nn.shape == (2, 0, 4, 3)
    d, i, _ = pcc.ops.knn(torch.empty(0, 5, 3, device=dev), torch.empty(0, 50, 3, device=dev), 4)
    assert d.shape == (0, 5, 4)
    assert pcc.ops.gather(torch.rand(2, 50, 7, device=dev), torch.empty(2, 0, dtype=torch.int64, device=dev)).shape == (2, 0, 7)
    _, bi = pcc.ops.ball_query(torch.empty(2, 0, 3, device=dev), p, 8, 0.2)
    assert bi.shape == (2, 0, 8)
    one = torch.rand(1, 1, 3, device=dev)
    d, i, _ = pcc.ops.knn(one, one, 1)
    assert float(d) == 0.0 and int(i) == 0
    r = pcc.ops.chamfer_forward(one, one)
    assert float(r["loss"]) == 0.0
    assert pcc.ops.fps(one, 1, None, pcc.ops.FLT_MAX).tolist() == [[0]]
    with pytest.raises(ValueError):
        pcc.ops.knn(torch.rand(2, 5, 3, device=dev), torch.rand(3, 50, 3, device=dev), 4)      # batch mismatch
    with pytest.raises(ValueError):
        pcc.ops.knn(torch.rand(2, 5, 2, device=dev), torch.rand(2, 50, 2, device=dev), 4)      # D != 3
    with pytest.raises(RuntimeError):
        pcc.ops.knn(torch.rand(2, 5, 3), torch.rand(2, 50, 3), 4)                              # CPU tensors: no fallback


@pytest.mark.parametrize("name", ["flat", "peaky", "wide"])
def test_pmf_to_cdf_reference_golden(pcc, golden_dir, name):
    """pcc_pmf_to_cdf_u16 against the reference's own pn_kit.pmf_to_cdf (tests/golden/ref_entropy.npz) + torchac's 16-bit
    conversion; pcc_cdf_to_u16 on the reference's float CDF; the rate estimate of pn_kit.estimate_bits_from_pmf."""
    g = np.load(os.path.join(golden_dir, "ref_entropy.npz"))
    cdf = torch.from_numpy(g[f"{name}_cdf"])
    L = cdf.shape[-1] - 1
    want = (cdf.mul(2 ** 16 - L).round().to(torch.int16) + torch.arange(L + 1, dtype=torch.int16)).numpy().view(np.uint16)
    assert np.array_equal(pcc.ops.pmf_to_cdf_u16(cu(g[f"{name}_pmf"])).cpu().numpy(), want)
    assert np.array_equal(pcc.ops.cdf_to_u16(cu(g[f"{name}_cdf"])).cpu().numpy(), want)
    from pcc_b200.train import estimate_bits_from_pmf
    bits = float(estimate_bits_from_pmf(cu(g[f"{name}_pmf"]), torch.from_numpy(g[f"{name}_sym"]).cuda()))
    assert abs(bits - float(g[f"{name}_bits"])) <= 1e-5 * float(g[f"{name}_bits"])


def test_scene_scale_cfg5(pcc, orc):
    """BASELINE cfg5 at full size: one 1,000,000-point S3DIS-shaped scene.  FPS and kNN patching are checked bit-exactly against
    the oracle on a prefix the oracle finishes in seconds (the first 2200 of the 7812 centres; 48 queries of the K = 256 search),
    and on the whole problem through properties: distinct in-range indices, ascending distances, the query's own point first."""
    scene = synth.scene_like(1_000_000, seed=3)
    xyz = cu(scene)
    start = np.array([12345], np.int64)
    S = 1_000_000 * 2 // 256                                         # compress.py:93 with ALPHA = 2, K = 256 -> 7812
    idx = pcc.ops.fps(xyz, S, cu(start), 1e10)
    got = idx.cpu().numpy()
    assert got.shape == (1, S) and len(np.unique(got)) == S and got.min() >= 0 and got.max() < 1_000_000
    assert np.array_equal(got[:, :2200], orc.fps(scene, 2200, start, 1e10, threads=8))   # FPS is a prefix-stable sequence; 2200 > one tag wrap
    centres = pcc.index_points(xyz, idx)
    d, i, nn = pcc.ops.knn(centres, xyz, 256, return_nn=True)
    dn, inn = d.cpu().numpy(), i.cpu().numpy()
    assert (np.diff(dn, axis=2) >= 0).all() and (dn[:, :, 0] == 0).all()             # sorted; a centre is a cloud point
    cen = centres.cpu().numpy()
    assert np.array_equal(scene[0][inn[0, :, 0]], cen[0])                              # nearest = the centre itself (or a duplicate)
    assert np.array_equal(nn.cpu().numpy()[0, ::500], scene[0][inn[0, ::500]])         # gathered neighbours = indexed points
    od, oi, _ = orc.knn_points(cen[:, :48], scene, 256, False, threads=8)
    assert np.array_equal(inn[:, :48], oi) and np.array_equal(dn[:, :48], od)
