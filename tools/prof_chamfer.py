"""Runs the Chamfer forward at the headline size (32 x 8192 x 8192) a few times (for ncu)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
import pcc_b200
from tools import synth
B = int(os.environ.get("B", 32))
x = torch.from_numpy(synth.modelnet_like(B, 8192, seed=1)).cuda()
y = torch.from_numpy(synth.decompressed_like(x.cpu().numpy())).cuda()
for _ in range(int(os.environ.get("ITERS", 3))):
    r = pcc_b200.ops.chamfer_forward(y, x, want_idx=False)
torch.cuda.synchronize()
print("ok", float(r["loss"]))
