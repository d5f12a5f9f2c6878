"""PPPF_AE forward (cfg3: 64 ShapeNet-shaped clouds x 2048 points): eager and graph replay."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
from pcc_b200 import graph as pgraph, pppf
from tools import synth
from tools.bench_ops import timeit
model = pppf.PPPF_AE(K=512, k=0, d=16, L=7)
model.load_state_dict(synth.seeded_module_state(model, 17))
model = model.cuda().eval()
sh = torch.from_numpy(synth.shapenet_like(64, 2048, seed=2)).cuda()
with torch.no_grad():
    b, m = timeit(lambda: model(sh), iters=10)
    print(f"PPPF_AE forward eager best {b:.3f} ms")
    replay = pgraph.capture(model, sh)
    b, m = timeit(lambda: replay(sh), iters=20)
    print(f"PPPF_AE forward graph best {b:.3f} ms")
