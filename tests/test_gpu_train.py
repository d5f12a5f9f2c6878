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
