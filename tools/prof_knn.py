"""In-patch kNN (2048 x 256 x 256, K = 16, neighbours only) once, for ncu."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
import pcc_b200
from tools import synth
xyz = torch.from_numpy(synth.modelnet_like(32, 8192, seed=1)).cuda()
start = torch.zeros(32, dtype=torch.int64, device="cuda")
cent = pcc_b200.index_points(xyz, pcc_b200.ops.fps(xyz, 64, start, 1e10))
_, _, patches = pcc_b200.ops.knn(cent, xyz, 256, True, True, 2.0)
patches = patches.reshape(32 * 64, 256, 3).contiguous()
for _ in range(2):
    out = pcc_b200.ops.knn(patches, patches, 16, return_nn=True, centre_sub=True, nn_only=True)
torch.cuda.synchronize()
print("ok")
