"""cfg5 legs: FPS 1M -> 7812 and kNN 7812 x 1M (K = 256), grid form vs brute force."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
import pcc_b200
from tools import synth
from tools.bench_ops import timeit
ops = pcc_b200.ops
sc = torch.from_numpy(synth.scene_like(1_000_000, seed=3)).cuda()
start = torch.zeros(1, dtype=torch.int64, device="cuda")
cen = pcc_b200.index_points(sc, ops.fps(sc, 7812, start, 1e10))
for name, g in (("grid", True), ("brute", False)):
    best, med = timeit(lambda: ops.knn(cen, sc, 256, True, True, grid=g), iters=5, warm=2)
    print(f"knn 7812 x 1M K=256 [{name}] best {best:.3f} ms median {med:.3f} ms", flush=True)
a = ops.knn(cen, sc, 256, True, True, grid=True); b = ops.knn(cen, sc, 256, True, True, grid=False)
print("identical:", all(torch.equal(x, y) for x, y in zip(a, b)))
for K in (16, 32):
    c512 = cen[:, :512].contiguous()
    for name, g in (("grid", True), ("brute", False)):
        best, med = timeit(lambda: ops.knn(c512, sc, K, grid=g), iters=5, warm=2)
        print(f"knn 512 x 1M K={K} [{name}] best {best:.3f} ms", flush=True)
if os.environ.get("FPS", "1") == "1":
    best, med = timeit(lambda: ops.fps(sc, 7812, start, 1e10), iters=3, warm=1)
    print(f"fps 1M -> 7812 best {best:.2f} ms")
