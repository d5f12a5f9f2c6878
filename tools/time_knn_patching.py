"""kNN patching (32 clouds x 64 centres x 8192 points, K = 256): CTA-per-query kernel vs the grid form."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
import pcc_b200
from tools import synth
from tools.bench_ops import timeit
ops = pcc_b200.ops
xyz = torch.from_numpy(synth.modelnet_like(32, 8192, seed=1)).cuda()
start = torch.zeros(32, dtype=torch.int64, device="cuda")
cen = pcc_b200.index_points(xyz, ops.fps(xyz, 64, start, 1e10))
for name, g in (("block", False), ("grid", True)):
    b, m = timeit(lambda: ops.knn(cen, xyz, 256, True, True, 2.0, nn_only=True, grid=g), iters=20)
    print(f"knn patching [{name}] best {b*1e3:.1f} us median {m*1e3:.1f} us", flush=True)
a = ops.knn(cen, xyz, 256, True, True, 2.0, grid=True); b = ops.knn(cen, xyz, 256, True, True, 2.0, grid=False)
print("identical:", all(torch.equal(x, y) for x, y in zip(a, b)))
