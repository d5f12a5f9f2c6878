"""CPU tests of the oracle: against the committed golden vectors (minted from the reference's own code, see
tests/golden/make_golden.py), against the live reference when /root/reference is present, and against an
independent torch statement of the PyTorch3D semantics."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from oracle import ref_loader
from tools import synth


@pytest.fixture(scope="module")
def g_ref(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_fps_gather.npz"))


@pytest.fixture(scope="module")
def g_p3d(golden_dir):
    return np.load(os.path.join(golden_dir, "p3d_ops.npz"))


@pytest.mark.parametrize("name", ["modelnet", "grid", "perm", "tiny"])
def test_fps_and_gather_match_reference_golden(g_ref, name):
    xyz = g_ref[f"{name}_xyz"]
    idx = orc.fps(xyz, int(g_ref[f"{name}_npoint"]), g_ref[f"{name}_start"], 1e10)
    assert np.array_equal(idx, g_ref[f"{name}_idx"])
    assert np.array_equal(orc.gather(xyz, idx), g_ref[f"{name}_gather"])
    assert np.array_equal(orc.gather(xyz, g_ref[f"{name}_idx3"]), g_ref[f"{name}_gather3"])


def test_synth_generators_are_stable(g_ref, g_p3d):
    assert np.array_equal(synth.modelnet_like(2, 8192, seed=11), g_ref["modelnet_xyz"])
    assert np.array_equal(synth.grid_quantised(3, 1000, depth=3, seed=12), g_ref["grid_xyz"])
    assert np.array_equal(synth.modelnet_like(2, 2048, seed=21), g_p3d["knn_patch_p"])


@pytest.mark.parametrize("name", ["knn_patch", "knn_dupq", "knn_ties", "knn_self", "knn_kgtn"])
def test_knn_golden(g_p3d, name):
    d, i, nn = orc.knn_points(g_p3d[f"{name}_q"], g_p3d[f"{name}_p"], int(g_p3d[f"{name}_K"]), True)
    assert np.array_equal(d, g_p3d[f"{name}_d"]) and np.array_equal(i, g_p3d[f"{name}_i"])
    assert np.array_equal(nn, g_p3d[f"{name}_nn"])


@pytest.mark.parametrize("name", ["ball_sa1", "ball_edge", "ball_few"])
def test_ball_query_golden(g_p3d, name):
    d, i = orc.ball_query(g_p3d[f"{name}_q"], g_p3d[f"{name}_p"], int(g_p3d[f"{name}_K"]), float(g_p3d[f"{name}_r"]))
    assert np.array_equal(i, g_p3d[f"{name}_i"]) and np.array_equal(d, g_p3d[f"{name}_d"])


@pytest.mark.parametrize("name", ["sfp", "sfp_pad", "sfp_ties"])
def test_sample_farthest_points_golden(g_p3d, name):
    pts, idx = orc.sample_farthest_points(g_p3d[f"{name}_x"], int(g_p3d[f"{name}_K"]))
    assert np.array_equal(idx, g_p3d[f"{name}_idx"]) and np.array_equal(pts, g_p3d[f"{name}_pts"])


def test_chamfer_golden_and_threads(g_p3d):
    x, y = g_p3d["cham_x"], g_p3d["cham_y"]
    for threads in (1, 4):
        loss, pc, dx, ix, dy, iy = orc.chamfer(x, y, threads=threads)
        assert loss == float(g_p3d["cham_loss"])
        assert np.array_equal(dx, g_p3d["cham_dx"]) and np.array_equal(iy, g_p3d["cham_iy"])
    gx, gy = orc.chamfer_bwd(x, y, ix, iy, 1.0)
    assert np.array_equal(gx, g_p3d["cham_gx"]) and np.array_equal(gy, g_p3d["cham_gy"])
    psnr, mse = orc.d1_psnr(x[0], y[0])
    assert psnr == float(g_p3d["d1_psnr"])


def test_knn_matches_torch_bruteforce_random():
    rng = np.random.default_rng(0)
    for trial in range(5):
        B, P1, P2, K = int(rng.integers(1, 4)), int(rng.integers(1, 40)), int(rng.integers(40, 300)), int(rng.integers(1, 40))
        q = rng.random((B, P1, 3), dtype=np.float32)
        p = synth.grid_quantised(B, P2, depth=2, seed=trial) if trial % 2 else rng.random((B, P2, 3), dtype=np.float32)
        d, i, _ = orc.knn_points(q, p, K)
        D = ((torch.from_numpy(q)[:, :, None] - torch.from_numpy(p)[:, None]) ** 2).sum(-1)
        ds, js = torch.sort(D, dim=2, stable=True)
        assert np.array_equal(ds[:, :, :K].numpy(), d) and np.array_equal(js[:, :, :K].numpy(), i)


def test_chamfer_backward_matches_autograd():
    rng = np.random.default_rng(1)
    x = torch.from_numpy(rng.random((2, 50, 3), dtype=np.float32)).requires_grad_()
    y = torch.from_numpy(rng.random((2, 70, 3), dtype=np.float32)).requires_grad_()
    D = ((x[:, :, None] - y[:, None]) ** 2).sum(-1)
    loss = (D.min(2)[0].mean(1) + D.min(1)[0].mean(1)).mean()
    loss.backward()
    l, _, _, ix, _, iy = orc.chamfer(x.detach().numpy(), y.detach().numpy())
    gx, gy = orc.chamfer_bwd(x.detach().numpy(), y.detach().numpy(), ix, iy, 1.0)
    assert abs(l - loss.item()) < 1e-6
    assert np.allclose(gx, x.grad.numpy(), atol=1e-7) and np.allclose(gy, y.grad.numpy(), atol=1e-7)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
def test_fps_matches_live_reference():
    pn = ref_loader.load("pn_kit")
    torch.manual_seed(3)
    for B, N, S in [(2, 500, 33), (1, 2048, 128), (3, 64, 64)]:
        xyz = torch.rand(B, N, 3)
        if N == 500:
            xyz = (torch.floor(xyz * 4) + 0.5) / 4
        state = torch.get_rng_state()
        ref = pn.farthest_point_sample_batch(xyz, S)
        torch.set_rng_state(state)
        start = torch.randint(0, N, (B,), dtype=torch.long)
        assert np.array_equal(orc.fps(xyz.numpy(), S, start.numpy(), 1e10), ref.numpy())
        assert np.array_equal(orc.gather(xyz.numpy(), ref.numpy()), pn.index_points(xyz, ref).numpy())


# ---- octree centre coding: the oracle against the REFERENCE's own coder (tests/golden/ref_octree.npz) -------------------
OCTREE_CASES = ["k256", "k128", "k1024", "s300", "dup", "grid", "one", "edge"]


@pytest.fixture(scope="module")
def g_oct(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_octree.npz"))


@pytest.mark.parametrize("name", OCTREE_CASES)
def test_octree_oracle_matches_reference_golden(g_oct, name):
    c, N, min_bpp = g_oct[f"{name}_c"], int(g_oct[f"{name}_N"]), float(g_oct[f"{name}_min_bpp"])
    codes, total, depths = orc.encode_sampled_np(c, 1, N, min_bpp)
    assert depths == g_oct[f"{name}_depth"].tolist()
    assert total == int(g_oct[f"{name}_nbits"].sum())
    for b, code in enumerate(codes):
        n = int(g_oct[f"{name}_nbits"][b])
        assert np.array_equal(code, g_oct[f"{name}_bits"][b, :n])
        assert np.array_equal(orc.bits_to_bytes(code), g_oct[f"{name}_bytes"][b, :(n + 7) // 8])
        assert np.array_equal(orc.octree_decode_ref(code), g_oct[f"{name}_rec"][b])
        assert np.array_equal(orc.octree_quantise(c[b], 1, depths[b])[1], g_oct[f"{name}_uniq{b}"])


def test_octree_oracle_fixed_depth_and_live_reference(g_oct):
    c = g_oct["k256_c"][0]
    for d in (1, 2, 5, 9):
        assert np.array_equal(orc.octree_encode(c, 1, d), g_oct[f"fixed_d{d}"])
    if not ref_loader.available():
        pytest.skip("/root/reference not present")
    pn = ref_loader.load("pn_kit")
    x = synth.uniform_cube(3, 48, seed=77) * np.float32(0.98) + np.float32(0.01)
    codes, bits = pn.encode_sampled_np(x, scale=1, N=2048, min_bpp=0.5)
    ocodes, obits, _ = orc.encode_sampled_np(x, 1, 2048, 0.5)
    assert bits == obits and all(np.array_equal(a, b) for a, b in zip(codes, ocodes))
    assert np.array_equal(pn.decode_sampled_np(codes, scale=1), np.stack([orc.octree_decode_ref(k) for k in ocodes]))


def test_eval_metrics_oracle_vs_reference_golden(golden_dir):
    """calc_uc (eval.py:127-151): the golden holds the output of the reference's own function; it takes distances from
    torch.cdist's matmul formulation (fp32), so the restatement (direct d2) agrees to the 5th digit -- 1e-4 relative."""
    g = np.load(os.path.join(golden_dir, "ref_eval.npz"))
    x = synth.modelnet_like(3, 8192, seed=81)
    y = synth.decompressed_like(x, seed=82)
    for b in range(2):
        assert abs(orc.calc_uc(x[b], y[b]) - g["uc"][b]) <= 1e-4 * g["uc"][b]
    psnr, mse = orc.p2plane_psnr(x[0], y[0])
    assert abs(psnr - g["p2plane"][0, 0]) < 1e-9 and psnr > orc.d1_psnr(x[0], y[0])[0]   # |diff . n| <= |diff|


# ---- entropy stage: pn_kit.pmf_to_cdf (run from the reference when present) + the restated torchac coder ----------------
@pytest.mark.parametrize("L,n,peaky", [(7, 1024, 1.0), (7, 20000, 6.0), (3, 100, 0.1), (16, 3000, 3.0), (7, 2000, 30.0)])
def test_range_coder_roundtrip_and_size(L, n, peaky):
    rng = np.random.default_rng(L * n)
    pmf = torch.softmax(torch.from_numpy(rng.normal(size=(n, L)).astype(np.float32) * peaky), -1)
    cdf = orc.pmf_to_cdf_u16(pmf.numpy())
    # pn_kit.pmf_to_cdf + torchac's _convert_to_int_and_normalize, stated with torch CPU ops
    c = torch.cat([torch.zeros(n, 1), pmf.cumsum(-1)], -1).clamp(max=1.0)
    if ref_loader.available():
        c = ref_loader.load("pn_kit").pmf_to_cdf(pmf)                       # the reference's own function
    ci = c.mul(2 ** 16 - L).round().to(torch.int16) + torch.arange(L + 1, dtype=torch.int16)
    assert np.array_equal(ci.numpy().view(np.uint16), cdf)
    assert (np.diff(cdf[:, :-1].astype(np.int64), axis=1) > 0).all()        # every symbol keeps a non-zero interval
    sym = np.array([rng.choice(L, p=p / p.sum()) for p in pmf.numpy()], np.int16)
    data = orc.range_encode(cdf, sym)
    assert np.array_equal(orc.range_decode(cdf, data), sym)
    ideal = -np.log2(np.maximum(pmf.numpy()[np.arange(n), sym], 2.0 ** -16)).sum() / 8
    assert len(data) <= ideal * 1.02 + 8                                    # an arithmetic coder sits within bits of the entropy


def test_range_coder_known_streams():
    """Hand-checkable streams of the 32-bit coder: a certain symbol costs nothing, a 1/2-probability symbol one bit."""
    cdf = np.array([[0, 32768, 0]], np.uint16).repeat(16, axis=0)            # two symbols, p = 1/2 each (last entry unused)
    bits = np.array([1, 0, 1, 1, 0, 0, 1, 0, 1, 1, 1, 0, 0, 1, 0, 1], np.int16)
    data = orc.range_encode(cdf, bits)
    assert len(data) == 3 and np.array_equal(np.unpackbits(np.frombuffer(data, np.uint8))[:16], bits)
    assert np.array_equal(orc.range_decode(cdf, data), bits)


@pytest.mark.parametrize("name", ["flat", "peaky", "wide"])
def test_pmf_to_cdf_matches_reference_golden(golden_dir, name):
    """The 16-bit CDF of the oracle equals torchac's conversion applied to the output of the reference's own pn_kit.pmf_to_cdf."""
    g = np.load(os.path.join(golden_dir, "ref_entropy.npz"))
    cdf = torch.from_numpy(g[f"{name}_cdf"])
    L = cdf.shape[-1] - 1
    want = (cdf.mul(2 ** 16 - L).round().to(torch.int16) + torch.arange(L + 1, dtype=torch.int16)).numpy().view(np.uint16)
    assert np.array_equal(orc.pmf_to_cdf_u16(g[f"{name}_pmf"]), want)


def test_torch_modules_match_reference_ae_golden(golden_dir):
    """oracle/torch_modules.py (the CPU baseline's network bodies) against the outputs of the reference's own AE.AE
    (tests/golden/ae_modules.npz, minted by make_golden.py::ae_modules from /root/reference): bit for bit."""
    from oracle import torch_modules as tm
    g = np.load(os.path.join(golden_dir, "ae_modules.npz"))
    sd = synth.seeded_state_dict(synth.ae_shapes(128, 16, 7), 11)
    x = torch.from_numpy(g["x"])
    with torch.no_grad():
        feat = tm.set_abstraction(sd, x.transpose(2, 1))
        latent, lq = tm.ae_encode(sd, x, L=7)
        rec = tm.ae_decode(sd, lq, k=128)
    assert np.array_equal(feat.numpy(), g["sa_feat"])
    assert np.array_equal(latent.numpy(), g["latent"]) and np.array_equal(lq.numpy(), g["lq"])
    assert np.array_equal(rec.numpy(), g["new_xyz"])


def _state_keys(path_first, stub_with_oracle):
    """state_dict keys / shapes of the network classes found under `path_first`, collected in a fresh interpreter."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = f"""
import json, sys
sys.path.insert(0, {root!r})
from oracle import ref_loader
ref_loader.install_stubs()
sys.path.insert(0, {path_first!r})
import AE, PPPF_AE, pppe_pcd_ae
out = {{}}
for name, m in (("AE.AE", AE.AE(256, 128, 16, 7)), ("AE.prob", AE.ConditionalProbabilityModel(7, 16)),
                ("pppe.encoder", pppe_pcd_ae.PointNet2EncoderFull(latent_dim=256)),
                ("PPPF_AE.PPPF_AE", PPPF_AE.PPPF_AE(K=512, k=0, d=16, L=7)), ("PPPF_AE.prob", PPPF_AE.ConditionalProbabilityModel(7, 16)),
                ("PPPF_AE.AE", PPPF_AE.AE(K=256, k=0, d=16, L=7))):
    out[name] = {{k: list(v.shape) for k, v in m.state_dict().items()}}
    out[name + ":attrs"] = sorted(k for k in vars(m) if not k.startswith("_"))
print(json.dumps(out))
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference is not present")
def test_standins_match_reference_state_dict_keys():
    """tests/standins/ (what the GPU box tests install() against) mirror the real reference's classes: same state_dict keys
    and shapes, same public instance attributes."""
    here = os.path.dirname(os.path.abspath(__file__))
    assert _state_keys(os.path.join(here, "standins"), True) == _state_keys(ref_loader.REFERENCE_ROOT, True)


def test_standin_forwards_reproduce_reference_goldens(golden_dir):
    """The stand-ins' eager forward bodies, run on CPU with the oracle ops, reproduce the outputs of the real reference's
    AE.AE, PPPF_AE.PPPF_AE and pppe_pcd_ae.PointNet2EncoderFull (the committed goldens) -- so "stand-in under install() == reference under install()"."""
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    code = f"""
import sys
import numpy as np, torch
sys.path.insert(0, {root!r})
from oracle import ref_loader
from tools import synth
ref_loader.install_stubs()
sys.path.insert(0, {os.path.join(here, "standins")!r})
import AE, PPPF_AE, pppe_pcd_ae
g = np.load({os.path.join(golden_dir, "pppe_modules.npz")!r})
m = pppe_pcd_ae.PointNet2EncoderFull(latent_dim=256); m.load_state_dict(synth.seeded_module_state(m, 23)); m.eval()
with torch.no_grad():
    torch.manual_seed(11)
    latent, pooled = m(torch.from_numpy(g["x"]))
assert np.abs(latent.numpy() - g["latent"]).max() < 1e-5 and np.abs(pooled.numpy() - g["pooled"]).max() < 1e-5
g = np.load({os.path.join(golden_dir, "ae_modules.npz")!r})
m = AE.AE(256, 128, 16, 7); m.load_state_dict(synth.seeded_state_dict(synth.ae_shapes(128, 16, 7), 11)); m.eval()
with torch.no_grad():
    new_xyz, latent, lq = m(torch.from_numpy(g["x"]))
assert np.array_equal(lq.numpy(), g["lq"]) and np.abs(latent.numpy() - g["latent"]).max() < 1e-5
assert np.abs(new_xyz.numpy() - g["new_xyz"]).max() < 1e-5
g = np.load({os.path.join(golden_dir, "pppf_modules.npz")!r})
m = PPPF_AE.PPPF_AE(K=512, k=0, d=16, L=7); m.load_state_dict(synth.seeded_module_state(m, 17)); m.eval()
with torch.no_grad():
    recon, latent, lq = m(torch.from_numpy(g["x"]))
assert np.array_equal(lq.numpy(), g["lq"]) and np.abs(latent.numpy() - g["latent"]).max() < 1e-4
assert np.abs(recon.numpy() - g["recon"]).max() < 1e-4
print("ok")
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-3000:]
