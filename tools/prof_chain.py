"""Runs the three fused MLP chains of the AE at the headline size once (for ncu)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200"), os.path.join(ROOT, "tests")]
from pcc_b200 import mlp_ops
from test_gpu_mlp import make_layers

BS = int(os.environ.get("BS", 2048))
sa = make_layers([3, 32, 64, 128], [True, True, True], 1)
pna = make_layers([131, 128, 256], [True, True], 2)
dec = make_layers([144, 128, 64, 32, 3], [True, True, True, False], 3)
g = torch.rand(BS * 256 * 16, 3, device="cuda") - 0.5
xyz = torch.rand(BS * 256, 3, device="cuda") - 0.5
lat = torch.randint(-3, 4, (BS, 16), device="cuda").float()
lin = (torch.rand(BS * 128, 128, device="cuda") - 0.5).bfloat16()
for it in range(int(os.environ.get("ITERS", 2))):
    feat = mlp_ops.fused_chain(g, sa, group=16, out_dtype=torch.bfloat16)
    h = mlp_ops.fused_chain([(feat, 1), (xyz, 1)], pna, out_dtype=torch.bfloat16)
    o = mlp_ops.fused_chain([(lin, 1), (lat, 128)], dec)
torch.cuda.synchronize()
print("ok", feat.shape, h.shape, o.shape)
