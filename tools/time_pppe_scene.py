import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
from pcc_b200 import pppe, ops
from tools import synth
from tools.bench_ops import timeit
xyz = torch.from_numpy(synth.scene_like(1_000_000, seed=3)).cuda()
enc = pppe.PointNet2EncoderFull(latent_dim=256)
enc.load_state_dict(synth.seeded_module_state(enc, 23))
enc = enc.cuda().eval()
torch.manual_seed(11)
for i in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pppe.compress(enc, xyz, latent_bins=7)
    torch.cuda.synchronize(); print(f"call {i}: {1e3*(time.perf_counter()-t0):.2f} ms", flush=True)
start = torch.zeros(1, dtype=torch.int64, device="cuda")
for s in (0, 5, 999_999, 123_456):
    st = torch.full((1,), s, dtype=torch.int64, device="cuda")
    b, m = timeit(lambda: ops.fps(xyz, 512, st, 1e10), iters=3, warm=1)
    print(f"fps 1M -> 512 start {s}: {b:.2f} ms")
