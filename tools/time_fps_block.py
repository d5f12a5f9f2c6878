"""FPS timings at the shapes of the in-step / PPPF / scene-head samplings (CUDA events around the C-ABI call, best and median of 9)."""
import sys, os, torch
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "point-cloud-compression_b200")]
from pcc_b200 import ops
from tools import synth
for (B, N, S) in [(32, 8192, 64), (64, 2048, 512), (1, 8192, 64), (1, 1000000, 512)]:
    x = torch.from_numpy(synth.modelnet_like(B, N, seed=1) if N <= 8192 else synth.scene_like(N, seed=3)).cuda()
    st = torch.zeros(B, dtype=torch.int64, device="cuda")
    for _ in range(3): ops.fps(x, S, st, 1e10)
    torch.cuda.synchronize()
    t = []
    for _ in range(9):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); ops.fps(x, S, st, 1e10); e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    t.sort()
    print(f"fps {B}x{N}->{S}: best {t[0]*1e3:.1f} us median {t[4]*1e3:.1f} us", flush=True)
