"""GPU module parity against golden outputs of the REFERENCE's own nn.Modules (AE.AE, PPPF_AE.PPPF_AE) run on CPU in
fp32 with the oracle ops injected (tests/golden/make_golden.py).  The weights are regenerated from a seed on both sides.
The device bodies use bf16 operands / fp32 accumulation, hence the stated tolerances (relative to each tensor's scale)."""
import os

import numpy as np
import pytest
import torch

from tools import synth

pytestmark = pytest.mark.gpu

FEATURE_RTOL = 2e-2   # per-layer features vs the fp32 reference (max abs error / max abs value)
COORD_ATOL = 1e-2     # decoded coordinates


@pytest.fixture(scope="module")
def pcc():
    import __graft_entry__  # noqa: F401
    import pcc_b200
    return pcc_b200


def rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


def test_ae_forward_vs_reference_module_golden(pcc, golden_dir):
    from pcc_b200.modules import AE
    g = np.load(os.path.join(golden_dir, "ae_modules.npz"))
    ae = AE(256, 128, 16, 7)
    ae.load_state_dict(synth.seeded_state_dict(synth.ae_shapes(128, 16, 7), 11))
    ae = ae.cuda().eval()
    x = torch.from_numpy(g["x"]).cuda()
    with torch.no_grad():
        new_xyz, latent, lq = ae(x)                                  # AE.forward, AE.py:34-55
        _, feat = ae.sa(x.transpose(2, 1))                            # pn_kit.SetAbstraction.forward signature
    assert rel(feat.float().cpu().numpy(), g["sa_feat"]) < FEATURE_RTOL
    assert np.abs(latent.cpu().numpy() - g["latent"]).max() < 5e-3
    same = lq.cpu().numpy() == g["lq"]
    assert same.mean() > 0.97
    # the decoder is judged on the reference's own symbols, whatever the encoder's symbols were (one flipped symbol of a
    # near-.5 latent must not switch the check off)
    with torch.no_grad():
        dec = ae.decode_patches(torch.from_numpy(g["lq"]).cuda())
    assert np.abs(dec.cpu().numpy() - g["new_xyz"]).max() < COORD_ATOL
    if same.all():
        assert np.abs(new_xyz.cpu().numpy() - g["new_xyz"]).max() < COORD_ATOL


def test_pppf_encoder_decoder_vs_reference_module_golden(pcc, golden_dir):
    from pcc_b200 import pppf
    g = np.load(os.path.join(golden_dir, "pppf_modules.npz"))
    model = pppf.PPPF_AE(K=512, k=0, d=16, L=7)
    model.load_state_dict(synth.seeded_module_state(model, 17))
    model = model.cuda().eval()
    torch.set_grad_enabled(False)   # the fused inference bodies (with autograd on, the modules run their differentiable form)
    try:
        x = torch.from_numpy(g["x"]).cuda()
        xyz1, f1 = model.encoder.sa1(x, None)
        assert np.array_equal(xyz1.cpu().numpy(), g["xyz1"])             # FPS (start 0) centres: exact
        assert rel(f1.cpu().numpy(), g["f1"]) < FEATURE_RTOL
        # feed the reference's own intermediate features forward so each stage is judged on its own
        xyz2, f2 = model.encoder.sa2(torch.from_numpy(g["xyz1"]).cuda(), torch.from_numpy(g["f1"]).cuda())
        assert np.array_equal(xyz2.cpu().numpy(), g["xyz2"])
        assert rel(f2.cpu().numpy(), g["f2"]) < FEATURE_RTOL
        xyz3, f3 = model.encoder.sa3(torch.from_numpy(g["xyz2"]).cuda(), torch.from_numpy(g["f2"]).cuda())
        assert np.array_equal(xyz3.cpu().numpy(), g["xyz3"])
        assert rel(f3.cpu().numpy(), g["f3"]) < FEATURE_RTOL
        recon, latent, lq = model(x)                                     # end to end
        assert recon.shape == (2, 256, 3) and latent.shape == (2, 1024) and lq.shape == (2, 16)
        assert np.abs(latent.cpu().numpy() - g["latent"]).max() < 0.1     # sigmoid-spread latent in [-3.4, 3.4]
        dec = model.decoder(model.dec_proj(torch.from_numpy(g["lq"]).cuda()))
        assert np.abs(dec.cpu().numpy() - g["recon"]).max() < COORD_ATOL * max(1.0, float(np.abs(g["recon"]).max()))
    finally:
        torch.set_grad_enabled(True)


def test_pppe_encoder_vs_reference_module_golden(pcc, golden_dir):
    """pcc_b200.pppe.PointNet2EncoderFull ("fast pppe_pcd_ae compress", cfg5) against the outputs of the REFERENCE's own
    pppe_pcd_ae.PointNet2EncoderFull: FPS centres exact (same CPU-RNG draws), per-level features and the latent within the
    bf16 tolerance; every level also judged on the reference's own inputs."""
    from pcc_b200 import pppe
    g = np.load(os.path.join(golden_dir, "pppe_modules.npz"))
    model = pppe.PointNet2EncoderFull(latent_dim=256)
    model.load_state_dict(synth.seeded_module_state(model, 23))
    model = model.cuda().eval()
    x = torch.from_numpy(g["x"]).cuda()
    with torch.no_grad():
        torch.manual_seed(11)
        xyz1, f1 = model.sa_modules[0](x, None)
        assert np.array_equal(xyz1.cpu().numpy(), g["xyz1"])
        assert rel(f1.cpu().numpy(), g["f1"]) < FEATURE_RTOL
        xyz2, f2 = model.sa_modules[1](torch.from_numpy(g["xyz1"]).cuda(), torch.from_numpy(g["f1"]).cuda())
        assert np.array_equal(xyz2.cpu().numpy(), g["xyz2"])             # third CPU-RNG draw of the sequence
        assert rel(f2.cpu().numpy(), g["f2"]) < FEATURE_RTOL
        torch.manual_seed(11)
        latent, pooled = model(x)
        assert rel(pooled.cpu().numpy(), g["pooled"]) < 3e-2
        assert rel(latent.cpu().numpy(), g["latent"]) < 3e-2
        q, cond = pppe.compress(model, x, latent_bins=7)
        assert q.shape == (2, 256) and float(q.min()) >= 0 and float(q.max()) <= 6 and torch.equal(q, q.round())
    with pytest.raises(NotImplementedError):
        model.train()(x)


def test_pppe_msg_level_batches_its_samplings_on_scene_sized_clouds(pcc, monkeypatch):
    """Above 196,608 points the MSG level runs its branches' farthest-point samplings as ONE batched call (same CPU-RNG draws in
    the same order): centres and features are bit-identical to the branch-by-branch route."""
    from pcc_b200 import bodies, pppe
    model = pppe.PointNet2EncoderFull(latent_dim=256)
    model.load_state_dict(synth.seeded_module_state(model, 23))
    msg = model.cuda().eval().sa_modules[0]
    assert hasattr(msg, "branches") and len(msg.branches) == 2
    x = torch.from_numpy(synth.scene_like(200_000, seed=2)).cuda()
    with torch.no_grad():
        torch.manual_seed(5)
        xyz_a, f_a = bodies.pppe_msg_points(msg, x)
        nxt_a = int(torch.randint(0, 1000, (1,)))
        monkeypatch.setattr(bodies, "_MSG_BATCHED_FPS_MIN_POINTS", 1 << 30)
        torch.manual_seed(5)
        xyz_b, f_b = bodies.pppe_msg_points(msg, x)
        nxt_b = int(torch.randint(0, 1000, (1,)))
    assert torch.equal(xyz_a, xyz_b) and torch.equal(f_a, f_b)
    assert nxt_a == nxt_b                                               # the CPU generator advanced by the same draws


def test_pointnet_ops_wrapper_matches_reference_names(pcc):
    """PointnetPPOps (pointnet_sa_module.py:8-34): argument order and return types."""
    xyz = torch.from_numpy(synth.shapenet_like(2, 512, seed=3)).cuda()
    idx = pcc.PointnetPPOps.furthest_point_sample(xyz, 64)
    assert idx.shape == (2, 64) and idx.dtype == torch.int64 and bool((idx[:, 0] == 0).all())
    new_xyz = pcc.index_points(xyz, idx)
    knn = pcc.PointnetPPOps.ball_query(0.3, 16, xyz, new_xyz)
    assert hasattr(knn, "idx") and knn.idx.shape == (2, 64, 16)
    grouped = pcc.PointnetPPOps.group_points(xyz, knn)
    assert grouped.shape == (2, 64, 16, 3)
    d, i = pcc.PointnetPPOps.knn_point(8, xyz, new_xyz)
    assert d.shape == (2, 64, 8) and bool((d[:, :, 0] == 0).all())
