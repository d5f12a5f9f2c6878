"""SetAbstraction stage at the bench shape (32 clouds x 64 patches x 256 points, K = 16): the indexed chain kernel alone, with a
checksum of its output so that builds / variants (PCC_SA_SLOTS=2|4, PCC_B200_LIB=...) can be compared bit for bit."""
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
import pcc_b200
from pcc_b200 import mlp_ops
from tools import synth
from tools.bench_ops import timeit

ops = pcc_b200.ops
B = 32
xyz = torch.from_numpy(synth.modelnet_like(B, 8192, seed=1)).cuda()
start = torch.zeros(B, dtype=torch.int64, device="cuda")
cent = pcc_b200.index_points(xyz, ops.fps(xyz, 64, start, 1e10))
_, _, patches = ops.knn(cent, xyz, 256, True, True, 2.0)
patches = patches.reshape(B * 64, 256, 3).contiguous()
g = torch.Generator().manual_seed(5)
layers = []
for cin, cout in ((3, 32), (32, 64), (64, 128)):
    layers.append(((torch.randn(cout, cin, generator=g) / cin ** 0.5).cuda(), (0.1 * torch.randn(cout, generator=g)).cuda(), True))
idx8 = ops.knn_patch_u8(patches, 16)
out = mlp_ops.sa_chain_indexed(patches, idx8, layers, out_dtype=torch.bfloat16)
torch.cuda.synchronize()
digest = hashlib.sha1(out.view(torch.int16).cpu().numpy().tobytes()).hexdigest()[:16]
best, med = timeit(lambda: mlp_ops.sa_chain_indexed(patches, idx8, layers, out_dtype=torch.bfloat16), iters=30, warm=5)
kb, km = timeit(lambda: ops.knn_patch_u8(patches, 16), iters=30, warm=5)
print(f"sa_chain_indexed lib={os.path.basename(pcc_b200._lib.LIB_PATH)} slots={os.environ.get('PCC_SA_SLOTS', 'default')} "
      f"best {best * 1e3:.1f} us median {med * 1e3:.1f} us  sha1 {digest}  | knn_patch_u8 best {kb * 1e3:.1f} median {km * 1e3:.1f} us",
      flush=True)
