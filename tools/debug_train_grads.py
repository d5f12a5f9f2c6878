"""Per-parameter gradient error of the kernel training body against the fp32 autograd body (development aid)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
from pcc_b200.modules import AE
from tools import synth
torch.manual_seed(3)
torch.backends.cuda.matmul.allow_tf32 = False
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
    return {n: p.grad.clone() for n, p in ae.named_parameters()}
g_ref = run(ae.forward_train_fp32)
g_k = run(ae.forward_train)
for n, g in g_ref.items():
    a, b = g_k[n].double(), g.double()
    print(f"{n:40s} rel {float((a-b).norm()/(b.norm()+1e-30)):.4f} cos {float((a*b).sum()/(a.norm()*b.norm()+1e-30)):.5f} |ref| {float(b.norm()):.3e} |k| {float(a.norm()):.3e}")
