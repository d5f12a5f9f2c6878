"""Prints the bf16-vs-fp32 deviations of the codec path (to set the stated tolerances with margin)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
from tools import synth
from oracle import torch_modules as tm
from pcc_b200.codec import PatchCodec
from pcc_b200.modules import AE
sd = synth.seeded_state_dict(synth.ae_shapes(128, 16, 7), 11)
ae = AE(256, 128, 16, 7); ae.load_state_dict(sd); ae = ae.cuda().eval()
codec = PatchCodec(ae)
clouds = synth.modelnet_like(2, 8192, seed=31)
x = torch.from_numpy(clouds).cuda(); start = torch.tensor([5, 9]).cuda()
c = codec.compress(x, start)
rec = codec.decompress(c["latent_q"], c["centres"], 8192, c["center"], c["longest"])
met = codec.evaluate(rec, x).cpu().numpy()
for b in range(2):
    ref = tm.compress_decompress_eval(sd, clouds[b], int(start[b]), threads=8)
    lat = c["latent"][b].cpu().numpy()
    print("latent max abs err", np.abs(lat - ref["latent"]).max(), "mean", np.abs(lat - ref["latent"]).mean(), "symbol match", (c["latent_q"][b].cpu().numpy() == ref["latent_q"]).mean())
    rec_b = codec.decompress(torch.from_numpy(ref["latent_q"])[None].cuda(), c["centres"][b:b+1], 8192, c["center"][b:b+1], c["longest"][b:b+1])
    print("rec max abs err (same symbols)", np.abs(rec_b[0].cpu().numpy() - ref["rec"]).max(), "chamfer", met[b, 0], ref["chamfer"], "psnr", met[b, 1], ref["d1_psnr"])
