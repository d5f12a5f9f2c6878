"""Scene-scale FPS (N > 8192) timings: the bucketed form (fps_bucket.cu, default from 65,536 points) and the co-resident
multi-CTA kernel (PCC_FPS_PATH=grid) on the S3DIS-shaped scene and on uniform volumes; best and median of 5, CUDA events."""
import sys, os, torch
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "point-cloud-compression_b200")]
from pcc_b200 import ops
from tools import synth
g = torch.Generator(device="cuda").manual_seed(1)
cases = [("scene", 1, 1000000, 7812), ("uniform", 1, 1000000, 7812), ("uniform", 4, 100000, 1024), ("scene", 1, 250000, 1953)]
for (kind, B, N, S) in cases:
    if kind == "scene":
        x = torch.from_numpy(synth.scene_like(N, seed=3)).cuda()
    else:
        x = torch.rand(B, N, 3, device="cuda", generator=g)
    st = torch.zeros(B, dtype=torch.int64, device="cuda")
    res = {}
    for path in ("bucket", "grid"):
        os.environ["PCC_FPS_PATH"] = path
        for _ in range(2): res[path] = ops.fps(x, S, st, 1e10)
        torch.cuda.synchronize()
        t = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); ops.fps(x, S, st, 1e10); e1.record(); torch.cuda.synchronize()
            t.append(e0.elapsed_time(e1))
        t.sort()
        print(f"fps[{path}] {kind} {B}x{N}->{S}: best {t[0]:.3f} ms, median {t[2]:.3f} ms, {t[0]*1e3/S:.2f} us/iter", flush=True)
    print("  identical:", bool(torch.equal(res["bucket"], res["grid"])), flush=True)
