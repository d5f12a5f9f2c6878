"""clock64 timeline of the indexed SetAbstraction kernel (CTA 0, slot 0): needs a library built with -DPCC_SA_TICKS
(make -C point-cloud-compression_b200 ticks; PCC_B200_LIB=$PWD/point-cloud-compression_b200/pcc_b200/libpcc_b200_ticks.so python
tools/time_sa_ticks.py) -- bring-up aid."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
import pcc_b200
from pcc_b200 import mlp_ops, _lib
from tools import synth
ops = pcc_b200.ops
lib = _lib.load()
dbg = lib.pcc_debug_ws_timing
dbg.argtypes = [ctypes.c_void_p]
B = 32
xyz = torch.from_numpy(synth.modelnet_like(B, 8192, seed=1)).cuda()
cent = pcc_b200.index_points(xyz, ops.fps(xyz, 64, torch.zeros(B, dtype=torch.int64, device="cuda"), 1e10))
patches = ops.knn(cent, xyz, 256, True, True, 2.0)[2].reshape(B * 64, 256, 3).contiguous()
g = torch.Generator().manual_seed(5)
layers = [((torch.randn(co, ci, generator=g) / ci ** 0.5).cuda(), (0.1 * torch.randn(co, generator=g)).cuda(), True)
          for ci, co in ((3, 32), (32, 64), (64, 128))]
idx8 = ops.knn_patch_u8(patches, 16)
for _ in range(2):
    mlp_ops.sa_chain_indexed(patches, idx8, layers, out_dtype=torch.bfloat16)
buf = torch.zeros(8192, dtype=torch.int64, device="cuda")
dbg(buf.data_ptr())
mlp_ops.sa_chain_indexed(patches, idx8, layers, out_dtype=torch.bfloat16)
torch.cuda.synchronize()
dbg(None)
t = buf.cpu().tolist()
epi, mma, m2 = t[:2000], t[2048:4048], t[4096:4596]
base = min(x for x in epi + mma if x)
print("epilogue thread 0: [top] gather issue + wait L1 | epi1 | fence + bar + issue L2 | finish gather | wait L2 | epi2 | loop | total")
names = ("gi+w1", "epi1", "bar+i2", "fg", "w2", "epi2", "loop")
tot = [0] * 8
n = 0
for tile in range(10, 90):
    x = epi[tile * 7:tile * 7 + 8]
    if not all(x):
        break
    d = [x[i + 1] - x[i] for i in range(7)]
    for i in range(7):
        tot[i] += d[i]
    tot[7] += x[7] - x[0]
    n += 1
    if tile < 16:
        print(f" tile {tile}: [{x[0]-base:7d}] " + "  ".join(f"{nm} {v:5d}" for nm, v in zip(names, d)) + f"  total {x[7]-x[0]:6d}")
print(f" mean over {n} tiles: " + "  ".join(f"{nm} {v / n:6.0f}" for nm, v in zip(names, tot)) + f"  total {tot[7] / n:6.0f}")
print("MMA warp lane 0: [top] take xyz | wait acc free | issue L1 | wait X1 consumed | layer 0 (L2 issued inside at +..) | total")
names = ("wf", "i1", "tx", "wx", "l0")
tot = [0] * 7
n = 0
for tile in range(10, 90):
    x = mma[tile * 6:tile * 6 + 7]
    if not all(x):
        break
    d = [x[i + 1] - x[i] for i in range(5)]
    for i in range(5):
        tot[i] += d[i]
    tot[5] += x[6] - x[0]
    tot[6] += m2[tile] - x[0]
    n += 1
    if tile < 16:
        print(f" tile {tile}: [{x[0]-base:7d}] " + "  ".join(f"{nm} {v:5d}" for nm, v in zip(names, d)) + f"  L2 issued at +{m2[tile]-x[0]:5d}  total {x[6]-x[0]:6d}")
print(f" mean over {n} tiles: " + "  ".join(f"{nm} {v / n:6.0f}" for nm, v in zip(names, tot)) + f"  L2 at +{tot[6] / n:6.0f}  total {tot[5] / n:6.0f}")
