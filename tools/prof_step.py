"""One compress -> decompress -> eval step at the headline size (32 x 8192, K=256), for ncu captures."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
from pcc_b200.codec import PatchCodec
from pcc_b200.modules import AE
from tools import synth

ae = AE(256, 128, 16, 7)
ae.load_state_dict(synth.seeded_state_dict(synth.ae_shapes(128, 16, 7), 11))
ae = ae.cuda().eval()
codec = PatchCodec(ae, centre_mode="coded")
xyz = torch.from_numpy(synth.modelnet_like(32, 8192, seed=1000)).cuda()
start = torch.zeros(32, dtype=torch.int64, device="cuda")
for _ in range(int(os.environ.get("ITERS", 2))):
    out = codec.roundtrip(xyz, start)
torch.cuda.synchronize()
print("ok", out[2][:2].tolist())
