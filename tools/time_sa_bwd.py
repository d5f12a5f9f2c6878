"""SetAbstraction stack in training at the cfg2 shape (32 clouds x 64 patches x 256 points, K = 16): the fused forward and the
fused backward kernel (csrc/sa_bwd.cu) against the unfused training kernels (fold_first + linear_train x 2 + groupmax, and their
backward)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
import pcc_b200
from pcc_b200 import mlp_ops, train_ops as T
from tools import synth
from tools.bench_ops import timeit
ops = pcc_b200.ops
B = 32
xyz = torch.from_numpy(synth.modelnet_like(B, 8192, seed=1)).cuda()
cent = pcc_b200.index_points(xyz, ops.fps(xyz, 64, torch.zeros(B, dtype=torch.int64, device="cuda"), 1e10))
patches = ops.knn(cent, xyz, 256, True, True, 2.0)[2].reshape(B * 64, 256, 3).contiguous()
g = torch.Generator().manual_seed(5)
ws = []
for ci, co in ((3, 32), (32, 64), (64, 128)):
    ws += [(torch.randn(co, ci, generator=g) / ci ** 0.5).cuda().requires_grad_(), (0.1 * torch.randn(co, generator=g)).cuda().requires_grad_()]
layers = [(ws[0], ws[1], True), (ws[2], ws[3], True), (ws[4], ws[5], True)]
G = torch.randn(B * 64 * 256, 128, device="cuda")
idx8 = ops.knn_patch_u8(patches, 16)
best, med = timeit(lambda: mlp_ops.sa_chain_indexed(patches, idx8, [(w.detach(), b.detach(), r) for w, b, r in layers]), iters=10)
print(f"fused forward            best {best * 1e3:8.1f} us", flush=True)
best, med = timeit(lambda: mlp_ops.sa_chain_indexed_bwd(patches, idx8, [t.detach() for t in ws], G), iters=10)
print(f"fused backward           best {best * 1e3:8.1f} us", flush=True)

def unfused():
    for w in ws:
        w.grad = None
    _, _, grouped = ops.knn(patches, patches, 16, return_nn=True, centre_sub=True, nn_only=True)
    x1 = T.fold_first_train(grouped.reshape(-1, 3), ws[0], ws[1])
    out = T.mlp_train(x1, layers[1:], group=16, mode="pool", x0_is_relu=True)
    out.backward(G)

def fused():
    for w in ws:
        w.grad = None
    T.sa_indexed_train(patches, ops.knn_patch_u8(patches, 16), layers).backward(G)

for name, fn in (("unfused fwd + bwd (with kNN)", unfused), ("fused fwd + bwd (with kNN)", fused)):
    best, med = timeit(fn, iters=5)
    print(f"{name:32s} best {best * 1e3:8.1f} us", flush=True)
