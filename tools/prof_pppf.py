"""PPPF_AE forward (cfg3: 64 ShapeNet-shaped 2048-point clouds) or one IPDAE train step (cfg2), for ncu launch lists."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
from tools import synth

what = sys.argv[1] if len(sys.argv) > 1 else "pppf"
iters = int(os.environ.get("ITERS", 2))
if what == "pppf":
    from pcc_b200 import pppf
    model = pppf.PPPF_AE(K=512, k=0, d=16, L=7)
    model.load_state_dict(synth.seeded_module_state(model, 17))
    model = model.cuda().eval()
    sh = torch.from_numpy(synth.shapenet_like(64, 2048, seed=2)).cuda()
    with torch.no_grad():
        for _ in range(iters):
            out = model(sh)
    torch.cuda.synchronize()
    print("ok", tuple(out[0].shape))
else:
    from pcc_b200.train import Trainer
    tr = Trainer(state_dict=synth.seeded_state_dict(synth.ae_shapes(128, 16, 7), 11))
    xyz = torch.from_numpy(synth.modelnet_like(32, 8192, seed=1)).cuda()
    start = torch.zeros(32, dtype=torch.int64, device="cuda")
    for _ in range(iters):
        r = tr.step(xyz, start)
    torch.cuda.synchronize()
    print("ok", float(r["loss"]))
