"""Chamfer (distance-only form) on real reconstructions of the headline workload: three calls, for ncu captures."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
import pcc_b200
from pcc_b200.codec import PatchCodec
from pcc_b200.modules import AE
from tools import synth
ae = AE(256, 128, 16, 7)
ae.load_state_dict(synth.seeded_state_dict(synth.ae_shapes(128, 16, 7), 11))
ae = ae.cuda().eval()
codec = PatchCodec(ae, centre_mode="coded")
xyz = torch.from_numpy(synth.modelnet_like(32, 8192, seed=1000)).cuda()
start = torch.zeros(32, dtype=torch.int64, device="cuda")
c = codec.compress(xyz, start)
rec = codec.decompress(c["latent_q"], c["centres"], 8192, c["center"], c["longest"])
for _ in range(3):
    pcc_b200.ops.chamfer_forward(rec, xyz, want_idx=False)
torch.cuda.synchronize()
