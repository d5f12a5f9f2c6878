"""Multi-CTA FPS (N > 8192) timings at 2 / 13 / 123 / 147 CTAs per cloud: best and median of 7, CUDA events."""
import sys, os, torch
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "point-cloud-compression_b200")]
from pcc_b200 import ops
g = torch.Generator(device="cuda").manual_seed(1)
for (B, N, S) in [(1, 1000000, 7812), (4, 100000, 1024), (1, 16384, 256), (1, 1200000, 2048)]:
    x = torch.rand(B, N, 3, device="cuda", generator=g)
    st = torch.zeros(B, dtype=torch.int64, device="cuda")
    for _ in range(2): ops.fps(x, S, st, 1e10)
    torch.cuda.synchronize()
    t = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); ops.fps(x, S, st, 1e10); e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    t.sort()
    print(f"fps {B}x{N}->{S}: best {t[0]:.3f} ms, median {t[3]:.3f} ms, {t[0]*1e3/S:.2f} us/iter", flush=True)
