"""GPU test of the training step (train.py:148-247 body): gradients flow through the Chamfer backward kernel, the STE
quantiser and the network bodies; the loss goes down on a fixed batch."""
import numpy as np
import pytest
import torch

from tools import synth

pytestmark = pytest.mark.gpu


def test_train_step_reduces_chamfer_loss():
    import __graft_entry__  # noqa: F401
    from pcc_b200.train import Trainer
    torch.manual_seed(11)
    tr = Trainer(K=256, k=128, d=16, L=7, lr=1e-3, state_dict=synth.seeded_state_dict(synth.ae_shapes(128, 16, 7), 11))
    x = torch.from_numpy(synth.modelnet_like(2, 2048, seed=61)).cuda()
    start = torch.zeros(2, dtype=torch.int64, device="cuda")
    losses = []
    for _ in range(8):
        out = tr.step(x, start)
        losses.append(float(out["loss"]))
        assert np.isfinite(losses[-1])
    grads = [p.grad for p in tr.ae.parameters()]
    assert all(g is not None and torch.isfinite(g).all() for g in grads)
    assert any(float(g.abs().max()) > 0 for g in grads)
    assert min(losses[-3:]) < losses[0]


def test_training_forward_matches_fused_inference_forward():
    """The differentiable fp32 body and the fused bf16 inference body implement the same AE.forward."""
    import __graft_entry__  # noqa: F401
    from pcc_b200.modules import AE
    ae = AE(256, 128, 16, 7)
    ae.load_state_dict(synth.seeded_state_dict(synth.ae_shapes(128, 16, 7), 11))
    ae = ae.cuda()
    x = torch.rand(4, 256, 3, device="cuda") - 0.5
    rec_t, lat_t, lq_t = ae(x)                 # autograd on -> training body
    with torch.no_grad():
        rec_i, lat_i, lq_i = ae(x)             # fused inference body
    assert rec_t.requires_grad and not rec_i.requires_grad
    assert float((lat_t - lat_i).abs().max()) < 5e-3
    assert float((lq_t == lq_i).float().mean()) > 0.97
    # decoders compared on identical symbols, unconditionally: the training body's decoder half re-run on the fused body's lq
    with torch.no_grad():
        lin = ae.inv_pool(lq_i).view(4, -1, ae.k)
        h = torch.cat((lin, lq_i.unsqueeze(-1).repeat((1, 1, ae.k))), dim=1).permute(0, 2, 1).reshape(4 * ae.k, -1)
        for w, b, relu in ae.inv_mlp.layers():
            h = torch.addmm(b, h, w.t())
            h = torch.relu(h) if relu else h
    assert float((h.view(4, ae.k, 3) - rec_i).abs().max()) < 1e-2


def _rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def _r16(t):
    """Round to bf16 and back (differentiable: the casts pass the gradient through) -- where the kernels store bf16."""
    return t.to(torch.bfloat16).float()


def _emu_stack(x0, layers, group=0, mode="f32"):
    """torch statement of train_ops.mlp_train's FORWARD arithmetic under autograd: bf16 operands (weights and every stored
    activation rounded to bf16), fp32 accumulation; the backward is torch's fp32 autograd of that forward."""
    h = _r16(x0)
    L = len(layers)
    for i, (w, b, relu) in enumerate(layers):
        h = h @ _r16(w).t() + (b if b is not None else 0)
        if relu:
            h = torch.relu(h)
        if i + 1 < L or mode != "f32":
            h = _r16(h)
    if mode == "pool":
        h = h.view(-1, group, h.shape[1]).max(dim=1)[0]
    return h


def _emu_ae_forward(ae, xyz):
    from pcc_b200 import ops
    BS, P, _ = xyz.shape
    with torch.no_grad():
        _, _, grouped = ops.knn(xyz, xyz, 16, return_nn=True, centre_sub=True, nn_only=True)
    sa = ae.sa.layers()
    # the fused SetAbstraction kernels (chain_ws.cu / sa_bwd.cu): conv0 in fp32 stored as bf16, conv1's bias as a bf16 operand, conv2's
    # accumulators stay fp32 through the max, its bias (fp32) and the ReLU come after the max
    x1 = _r16(torch.relu(grouped.reshape(-1, 3) @ sa[0][0].t() + sa[0][1]))
    x2 = _r16(torch.relu(x1 @ _r16(sa[1][0]).t() + _r16(sa[1][1])))
    feat = torch.relu((x2 @ _r16(sa[2][0]).t()).view(-1, 16, sa[2][0].shape[0]).max(dim=1)[0] + sa[2][1])
    raw = _emu_stack(torch.cat((xyz.reshape(-1, 3), feat), dim=1), ae.pn.layers(), P, "pool")
    spread = ae.L - 0.2
    latent = torch.sigmoid(raw) * spread - spread / 2
    lq = ae.quantize(latent)
    lin = _emu_stack(lq, [(ae.inv_pool[i].weight, ae.inv_pool[i].bias, True) for i in (0, 2, 4)], mode="bf16").view(BS, -1, ae.k)
    x = torch.cat((lin, lq.unsqueeze(-1).repeat((1, 1, ae.k))), dim=1).permute(0, 2, 1).reshape(BS * ae.k, -1)
    return _emu_stack(x, ae.inv_mlp.layers()).view(BS, ae.k, 3), latent, lq


def test_kernel_training_gradients_match_the_autograd_bodies():
    """AE.forward_train (every forward / backward contraction on the pcc kernels: streamed tcgen05 GEMM, MN-major weight-gradient
    kernel, pooling kernels) against two torch autograd statements of the same body, same weights / patches / loss:
      (i) the bf16-operand model of its own arithmetic (_emu_ae_forward: same rounding points, so the same ReLU / arg-max
          patterns; torch's backward keeps fp32 gradients where the kernels store bf16): every parameter gradient within 1e-2
          relative L2 (measured: <= 6e-3);
      (ii) the plain fp32 body AE.forward_train_fp32 (the reference's default arithmetic): outputs and loss within the bf16
          tolerance, gradients with cosine > 0.8 -- a bf16 network flips ~1 % of the ReLU units and pooling winners whose
          pre-activations are near ties, which is what separates ANY reduced-precision training from fp32 gradients at depth."""
    import __graft_entry__  # noqa: F401
    from pcc_b200.modules import AE, ConditionalProbabilityModel
    torch.manual_seed(3)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ae = AE(256, 128, 16, 7)
        ae.load_state_dict(synth.seeded_state_dict(synth.ae_shapes(128, 16, 7), 11))
        ae = ae.cuda().train()
        xyz = torch.from_numpy(synth.modelnet_like(1, 2048, seed=9)).cuda().view(8, 256, 3) - 0.5
        target = torch.rand(8, 128, 3, device="cuda") - 0.5

        def run(fwd):
            ae.zero_grad(set_to_none=True)
            out, latent, lq = fwd(xyz)
            loss = (out - target).square().mean() + 1e-2 * latent.square().mean()
            loss.backward()
            return out.detach(), latent.detach(), float(loss), {n: p.grad.clone() for n, p in ae.named_parameters()}

        o_ref, l_ref, loss_ref, g_ref = run(ae.forward_train_fp32)
        o_emu, l_emu, loss_emu, g_emu = run(lambda x: _emu_ae_forward(ae, x))
        lib = __import__("pcc_b200")._lib.load()
        n0 = lib.pcc_launch_count()
        o_k, l_k, loss_k, g_k = run(ae.forward_train)
        assert lib.pcc_launch_count() - n0 >= 40            # 14 forward layers + pooling + 14 wgrad + 11 dgrad ...
        assert float((l_k - l_emu).abs().max()) < 2e-3 and float((o_k - o_emu).abs().max()) < 2e-3     # same arithmetic
        assert float((l_k - l_ref).abs().max()) < 2e-2 and float((o_k - o_ref).abs().max()) < 2e-2     # bf16 vs fp32
        assert abs(loss_k - loss_ref) < 2e-2 * abs(loss_ref)
        report = {}
        for name, g in g_emu.items():
            assert g_k[name].shape == g.shape and torch.isfinite(g_k[name]).all(), name
            if float(g.norm()) < 1e-12:
                continue
            cos = float((g_k[name].double() * g_ref[name].double()).sum() /
                        (g_k[name].double().norm() * g_ref[name].double().norm() + 1e-30))
            report[name] = (_rel_l2(g_k[name], g), cos)
        print({k: (round(a, 4), round(c, 4)) for k, (a, c) in report.items()})
        for name, (r, cos) in report.items():
            assert r < 1e-2 and cos > 0.8, (name, r, cos)
        # the probability model (PointNet + 3-layer head, softmax) against its fp32 body: shallow, so the comparison is direct
        prob = ConditionalProbabilityModel(7, 16).cuda().train()
        centres = torch.rand(3, 64, 3, device="cuda")
        sym = torch.randint(0, 7, (3, 64, 16, 1), device="cuda")

        def run_p(fwd):
            prob.zero_grad(set_to_none=True)
            pmf = fwd(centres)
            loss = -torch.log(pmf.gather(3, sym).clamp(min=1e-6)).mean()
            loss.backward()
            return pmf.detach(), {n: p.grad.clone() for n, p in prob.named_parameters()}

        p_ref, gp_ref = run_p(prob.forward_train_fp32)
        p_k, gp_k = run_p(prob.forward_train)
        assert float((p_k - p_ref).abs().max()) < 1e-2
        for name, g in gp_ref.items():
            if float(g.norm()) > 1e-12:
                assert _rel_l2(gp_k[name], g) < 0.15, (name, _rel_l2(gp_k[name], g))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def test_pppf_training_pass_and_fused_inference_agree():
    """PPPF_AE (PointNet++ SA x3 + FoldingNet, cfg3): with autograd on, the differentiable body runs (pcc kernels for sampling /
    ball query / grouping / Chamfer, torch layers for the MLPs); in eval mode it computes what the fused inference path computes,
    and a few optimisation steps reduce the Chamfer loss."""
    import __graft_entry__  # noqa: F401
    from pcc_b200 import pppf
    from pcc_b200.pytorch3d_compat import chamfer_distance
    torch.manual_seed(3)
    model = pppf.PPPF_AE(K=512, k=0, d=16, L=7)
    model.load_state_dict(synth.seeded_module_state(model, 17))
    model = model.cuda()
    x = torch.from_numpy(synth.shapenet_like(4, 2048, seed=71)).cuda()
    model.eval()
    rec_t, lat_t, lq_t = model(x)                          # autograd on -> differentiable body (BatchNorm in eval mode)
    with torch.no_grad():
        rec_i, lat_i, lq_i = model(x)                      # fused inference body
    assert rec_t.requires_grad and not rec_i.requires_grad
    assert float((lat_t - lat_i).abs().max()) < 2e-2      # bf16 tensor-core chain vs fp32 torch layers on a 1024-wide latent
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    losses = []
    for _ in range(6):
        opt.zero_grad(set_to_none=True)
        rec, _, _ = model(x)
        loss, _ = chamfer_distance(rec, x)                 # PPPF_AE.get_loss, PPPF_AE.py:168
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert all(np.isfinite(losses)) and min(losses[-2:]) < losses[0]
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())


def test_trainer_ddp_world_size_one_nccl():
    """Trainer(ddp=True) (train.py:148-247 as a DistributedDataParallel step): a 1-rank NCCL group exercises the wrappers, the
    bucketed gradient all-reduce and the optimiser over both models; the step must give the same losses as ddp=False from the
    same seed (a 1-rank all-reduce is the identity).  bench.py runs the same class at world sizes 2 / 4 / 8 (configs.cfg2)."""
    import os
    import torch.distributed as dist
    import __graft_entry__  # noqa: F401
    from pcc_b200.train import Trainer
    sd = synth.seeded_state_dict(synth.ae_shapes(128, 16, 7), 11)
    x = torch.from_numpy(synth.modelnet_like(2, 2048, seed=63)).cuda()
    start = torch.zeros(2, dtype=torch.int64, device="cuda")

    def run(ddp):
        torch.manual_seed(5)
        tr = Trainer(K=256, k=128, d=16, L=7, lr=1e-3, state_dict=sd, ddp=ddp)
        return [float(tr.step(x, start)["loss"]) for _ in range(4)], tr

    plain, _ = run(False)
    created = not dist.is_initialized()
    if created:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29611")
        dist.init_process_group("nccl", rank=0, world_size=1)
    try:
        ddp_losses, tr = run(True)
        assert all(p.grad is not None for p in tr.ae.parameters()) and all(p.grad is not None for p in tr.prob.parameters())
    finally:
        if created:
            dist.destroy_process_group()
    assert all(np.isfinite(ddp_losses)) and np.allclose(ddp_losses, plain, rtol=1e-4)


@pytest.mark.parametrize("BS,P", [(24, 256), (3, 200), (1, 16)])
def test_sa_fused_backward_matches_autograd_of_the_bf16_model(BS, P):
    """train_ops.sa_indexed_train: forward = the inference SetAbstraction kernel, backward = csrc/sa_bwd.cu (activations recomputed
    per tile, weight gradients accumulated in TMEM) against torch autograd on the bf16-operand statement of the same forward
    (pn_kit.py:190-207): outputs within bf16 rounding, every parameter gradient within 1e-2 relative L2 (the kernel stores dY2,
    dX2 and dX1 as bf16 where torch keeps fp32).  24 x 256: 768 tiles, several per slot and ragged over the CTAs; 3 x 200: a patch
    size that is not a power of two (75 tiles: one slot of the last CTA stays empty); 1 x 16: the smallest legal shape (one patch
    of K points, two tiles, one CTA)."""
    import __graft_entry__  # noqa: F401
    from pcc_b200 import ops, train_ops
    torch.manual_seed(5)
    patches = (torch.from_numpy(synth.modelnet_like(1, BS * P, seed=21)).cuda().view(BS, P, 3) - 0.5).contiguous()
    idx8 = ops.knn_patch_u8(patches, 16)
    g = torch.Generator().manual_seed(7)
    ws = []
    for cin, cout in ((3, 32), (32, 64), (64, 128)):
        ws.append((torch.randn(cout, cin, generator=g) / cin ** 0.5).cuda().requires_grad_())
        ws.append((0.1 * torch.randn(cout, generator=g)).cuda().requires_grad_())
    G = torch.randn(BS * P, 128, generator=g).cuda()

    out = train_ops.sa_indexed_train(patches, idx8, [(ws[0], ws[1], True), (ws[2], ws[3], True), (ws[4], ws[5], True)])
    (out * G).sum().backward()
    got = [w.grad.clone() for w in ws]
    for w in ws:
        w.grad = None

    nb = torch.gather(patches, 1, idx8.long().reshape(BS, P * 16, 1).expand(-1, -1, 3)).view(BS, P, 16, 3)
    local = (nb - patches.unsqueeze(2)).reshape(-1, 3)                                    # pn_kit.py:191
    x1 = _r16(torch.relu(local @ ws[0].t() + ws[1]))
    x2 = _r16(torch.relu(x1 @ _r16(ws[2]).t() + _r16(ws[3])))
    y2 = x2 @ _r16(ws[4]).t()
    ref = torch.relu(y2.view(-1, 16, 128).max(dim=1)[0] + ws[5])
    (ref * G).sum().backward()
    assert (out - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
    for name, a, w in zip(("w0", "b0", "w1", "b1", "w2", "b2"), got, ws):
        r = (a - w.grad).norm().item() / w.grad.norm().item()
        assert r < 1e-2, f"{name}: relative L2 error {r:.3e}"

    # the gradient as autograd hands it over in the AE: a bf16 column slice of the PointNet stack's [M, 192] input gradient, read in
    # place -- the same six tensors as from an fp32 contiguous copy of those values
    from pcc_b200 import mlp_ops
    wide = torch.zeros(BS * P, 192, dtype=torch.bfloat16, device="cuda")
    wide[:, :128] = G.to(torch.bfloat16)
    params = [w.detach() for w in ws]
    in_place = mlp_ops.sa_chain_indexed_bwd(patches, idx8, params, wide[:, :128])
    copied = mlp_ops.sa_chain_indexed_bwd(patches, idx8, params, wide[:, :128].float().contiguous())
    for a, b in zip(in_place, copied):   # the CTAs add their partial sums with fp32 atomics: the order, not the values, differs
        assert (a - b).norm().item() <= 1e-5 * b.norm().item()


def test_inference_after_fused_optimiser_steps_sees_the_new_weights():
    """torch's fused Adam kernel updates parameters without bumping their version counters, which every weight cache keys on:
    the trainer invalidates them (mlp_ops.invalidate_weight_caches).  The fused inference forward of the SAME module, and a
    captured round-trip graph, must follow the trained weights."""
    import __graft_entry__  # noqa: F401
    from pcc_b200.codec import PatchCodec
    from pcc_b200.train import Trainer
    torch.manual_seed(2)
    tr = Trainer(K=256, k=128, d=16, L=7, lr=1e-2, state_dict=synth.seeded_state_dict(synth.ae_shapes(128, 16, 7), 11))
    x = torch.from_numpy(synth.modelnet_like(2, 2048, seed=61)).cuda()
    start = torch.zeros(2, dtype=torch.int64, device="cuda")
    codec = PatchCodec(tr.ae.eval(), centre_mode="coded")
    run = codec.graphed_roundtrip(2, 2048)
    # [2] = the per-cloud metrics (Chamfer, D1 PSNR, MSE) of the reconstruction: they move with every weight of the decoder
    before, before_g = codec.roundtrip(x, start)[2].clone(), run(x, start)[2].clone()
    assert torch.equal(before, before_g)
    for _ in range(3):
        tr.step(x, start)
    tr.ae.eval()
    with pytest.raises(RuntimeError):                          # the old capture points at the old packed weights
        run(x, start)
    run = codec.graphed_roundtrip(2, 2048)
    after, after_g = codec.roundtrip(x, start)[2].clone(), run(x, start)[2].clone()
    assert torch.equal(after, after_g)
    assert not torch.equal(after, before)                      # three steps at lr = 1e-2 move the reconstruction
    fresh = PatchCodec(tr.ae, centre_mode="coded")             # nothing cached in this object
    assert torch.equal(fresh.roundtrip(x, start)[2], after)
