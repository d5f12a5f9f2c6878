"""Scene-scale FPS (N > 8192) timings: the default route (a head of co-resident iterations, then the bucketed form of
fps_bucket.cu), its A/B switches and the co-resident multi-CTA kernel alone (PCC_FPS_PATH=grid) on the S3DIS-shaped scene and on
uniform volumes; best and median of 5, CUDA events.  Variant syntax: path[:head=k][:morton]."""
import sys, os, torch
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "point-cloud-compression_b200")]
from pcc_b200 import ops
from tools import synth
g = torch.Generator(device="cuda").manual_seed(1)
cases = [("scene", 1, 1000000, 7812), ("uniform", 1, 1000000, 7812), ("scene", 1, 1000000, 512), ("uniform", 4, 100000, 1024),
         ("scene", 1, 250000, 1953)]
variants = ("bucket", "bucket:morton", "bucket:head=0", "bucket:head=512", "grid")
for (kind, B, N, S) in cases:
    if kind == "scene":
        x = torch.from_numpy(synth.scene_like(N, seed=3)).cuda()
    else:
        x = torch.rand(B, N, 3, device="cuda", generator=g)
    st = torch.zeros(B, dtype=torch.int64, device="cuda")
    res = {}
    for var in variants:
        f = var.split(":")
        for k in ("PCC_FPS_HEAD", "PCC_FPS_CURVE"):
            os.environ.pop(k, None)
        os.environ["PCC_FPS_PATH"] = f[0]
        for o in f[1:]:
            if o.startswith("head="):
                os.environ["PCC_FPS_HEAD"] = o[5:]       # iterations of the co-resident kernel before the hand-over
            elif o == "morton":
                os.environ["PCC_FPS_CURVE"] = "morton"    # Z-curve cell order instead of the Hilbert curve
        for _ in range(2): res[var] = ops.fps(x, S, st, 1e10)
        torch.cuda.synchronize()
        t = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); ops.fps(x, S, st, 1e10); e1.record(); torch.cuda.synchronize()
            t.append(e0.elapsed_time(e1))
        t.sort()
        print(f"fps[{var}] {kind} {B}x{N}->{S}: best {t[0]:.3f} ms, median {t[2]:.3f} ms, {t[0]*1e3/S:.2f} us/iter", flush=True)
    print("  identical:", all(bool(torch.equal(res[v], res["grid"])) for v in variants), flush=True)
