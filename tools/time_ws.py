"""clock64 ticks of the warp-specialised SA chain (CTA 0): MMA thread and epilogue thread 0 -- bring-up aid."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200"), os.path.join(ROOT, "tests")]
from pcc_b200 import mlp_ops, _lib
from test_gpu_mlp import make_layers
lib = _lib.load()
dbg = lib.pcc_debug_ws_timing
dbg.argtypes = [ctypes.c_void_p]
BS = 2048
sa = make_layers([3, 32, 64, 128], [True, True, True], 1)
g = torch.rand(BS * 256 * 16, 3, device="cuda") - 0.5
for _ in range(2):
    mlp_ops.fused_chain(g, sa, group=16, out_dtype=torch.bfloat16)
buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
dbg(buf.data_ptr())
mlp_ops.fused_chain(g, sa, group=16, out_dtype=torch.bfloat16)
torch.cuda.synchronize()
dbg(None)
t = buf.cpu().tolist()
epi, mma = t[:500], t[512:1012]
base = min(x for x in epi + mma if x)
# epilogue thread 0 of slot 0: 7 ticks per tile: start | L0+arrive | wait1 | epi1 | arrive | wait2 | epi2
print("epilogue (slot 0): [start] L0 | wait1 | epi1 | arrive | wait2 | epi2 | total")
for tile in range(2, 8):
    x = epi[tile * 7:tile * 7 + 8]
    d = [x[i + 1] - x[i] for i in range(6)]
    print(f" tile {tile}: [{x[0]-base:6d}] L0 {d[0]:5d}  w1 {d[1]:5d}  e1 {d[2]:5d}  a {d[3]:5d}  w2 {d[4]:5d}  e2 {d[5]:5d}  total {x[7]-x[0]:6d}")
print("mma warp (slot 0): per tile [start-wait] wait1 | issue1 | wait2 | issue2")
for tile in range(2, 8):
    x = mma[tile * 4:tile * 4 + 5]
    prev_end = mma[tile * 4 - 1]
    print(f" tile {tile}: [{prev_end-base:6d}] w1 {x[0]-prev_end:5d} i1 {x[1]-x[0]:4d} w2 {x[2]-x[1]:5d} i2 {x[3]-x[2]:4d}")
