"""PointNet of the AE (2048 patches x 256 positions): fused kernel vs front chain + tail."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
from pcc_b200 import mlp_ops
from tools.bench_ops import timeit
torch.manual_seed(1)
M = 2048 * 256
feat = torch.relu(torch.randn(M, 128, device="cuda")).to(torch.bfloat16)
xyz = (torch.rand(M, 3, device="cuda") - 0.5)
dims = [131, 128, 256, 512, 16]
layers = [(torch.randn(o, i, device="cuda") / i ** 0.5, torch.randn(o, device="cuda") * 0.1, l < 3) for l, (i, o) in enumerate(zip(dims[:-1], dims[1:]))]
b, m = timeit(lambda: mlp_ops.pointnet_fused(feat, xyz, layers), iters=20)
print(f"pointnet fused: best {b*1e3:.1f} us median {m*1e3:.1f} us  ({2*M*(131*128+128*256+256*512+512*16)/b/1e9:.0f} TFLOP/s)")
b, m = timeit(lambda: mlp_ops.run_chain([(feat, 1), (xyz, 1)], layers, group=256), iters=20)
print(f"front chain + tail: best {b*1e3:.1f} us median {m*1e3:.1f} us")
