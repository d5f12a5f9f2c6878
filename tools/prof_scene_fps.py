"""cfg5 scene FPS (1,000,000 points -> 7812 centres) twice: warm-up, then the launch an ncu capture of fps_bucket_kernel takes."""
import sys, os, torch
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "point-cloud-compression_b200")]
from pcc_b200 import ops
from tools import synth
x = torch.from_numpy(synth.scene_like(1_000_000, seed=3)).cuda()
st = torch.tensor([12345], dtype=torch.int64, device="cuda")
for _ in range(2):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); ops.fps(x, 7812, st, 1e10); e1.record(); torch.cuda.synchronize()
    print("fps 1M -> 7812: %.3f ms" % e0.elapsed_time(e1), flush=True)
