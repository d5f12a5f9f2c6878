"""In-patch kNN (2048 patches x 256 x 256, K = 16) timing: filter form vs the old form (PCC_KNN_THREAD_OLD=1), nn_only like the
SetAbstraction body calls it, and with every output."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
import pcc_b200
from tools import synth
from tools.bench_ops import timeit
ops = pcc_b200.ops
B = 32
xyz = torch.from_numpy(synth.modelnet_like(B, 8192, seed=1)).cuda()
start = torch.zeros(B, dtype=torch.int64, device="cuda")
cent = pcc_b200.index_points(xyz, ops.fps(xyz, 64, start, 1e10))
_, _, patches = ops.knn(cent, xyz, 256, True, True, 2.0)
patches = patches.reshape(B * 64, 256, 3)
for name, fn in (("nn_only", lambda: ops.knn(patches, patches, 16, True, True, nn_only=True)),
                 ("all outputs", lambda: ops.knn(patches, patches, 16, True, True)),
                 ("idx only", lambda: ops.knn(patches, patches, 16))):
    best, med = timeit(fn, iters=20)
    print(f"knn in-patch {B*64}x256x256 K16 [{name}] best {best*1e3:.1f} us median {med*1e3:.1f} us", flush=True)
