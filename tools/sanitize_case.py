"""One small pass over every kernel family of libpcc_b200.so, for `compute-sanitizer --tool {memcheck,racecheck,synccheck}`
(one tool per gpurun call, B200_PROFILING.md).  Sizes are the smallest that still reach every kernel: the IPDAE
round trip of 2 clouds (block FPS, octree coder, block / thread kNN, the warp-specialised tcgen05 chains, PointNet tail, streamed
GEMMs, assemble, grid Chamfer, eval), the multi-CTA FPS and warp kNN on a 20 k-point cloud, ball query, brute-force Chamfer +
backward, PPPF_AE and the pppe encoder on 2 clouds, the entropy stage."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
import pcc_b200  # noqa: E402
from pcc_b200 import ops, pppe, pppf  # noqa: E402
from pcc_b200.codec import PatchCodec  # noqa: E402
from pcc_b200.modules import AE, ConditionalProbabilityModel  # noqa: E402
from tools import synth  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(11)
with torch.no_grad():
    ae = AE(256, 128, 16, 7)
    ae.load_state_dict(synth.seeded_state_dict(synth.ae_shapes(128, 16, 7), 11))
    codec = PatchCodec(ae.to(dev).eval(), centre_mode="coded")
    clouds = torch.from_numpy(synth.modelnet_like(2, 8192, seed=6)).to(dev)
    start = torch.zeros(2, dtype=torch.int64, device=dev)
    lat_q, cen, met, rec, octree = codec.roundtrip(clouds, start, return_octree=True)
    prob = ConditionalProbabilityModel(7, 16).to(dev).eval()
    data, nbytes = codec.encode_latents(prob, lat_q.float(), cen)
    back = codec.decode_latents(prob, cen, data, nbytes)
    assert torch.equal(back, lat_q.float())
    print("round trip + entropy stage ok", met[:, 1].tolist())
    big = torch.from_numpy(synth.scene_like(20000, seed=3)).to(dev)
    idx = ops.fps(big, 300, start[:1], 1e10)                       # multi-CTA cooperative FPS
    q = ops.gather(big, idx)
    ops.knn(q, big, 64, return_nn=True, centre_sub=True)            # warp-per-query kNN
    sh = torch.from_numpy(synth.shapenet_like(2, 2048, seed=2)).to(dev)
    c = ops.gather(sh, ops.fps(sh, 128, None, ops.FLT_MAX))
    ops.ball_query(c, sh, 32, 0.2)
    small = torch.rand(3, 500, 3, device=dev)
    r = ops.chamfer_forward(small, torch.rand(3, 700, 3, device=dev))   # brute-force path
    print("fps grid / knn warp / ball / brute chamfer ok", float(r["per_cloud"].sum()))
    model = pppf.PPPF_AE(K=512, k=0, d=16, L=7)
    model.load_state_dict(synth.seeded_module_state(model, 17))
    out = model.to(dev).eval()(sh)
    enc = pppe.PointNet2EncoderFull(latent_dim=256)
    enc.load_state_dict(synth.seeded_module_state(enc, 23))
    lq, _ = pppe.compress(enc.to(dev).eval(), sh)
    print("pppf / pppe ok", tuple(out[0].shape), tuple(lq.shape))
x = torch.rand(2, 1500, 3, device=dev, requires_grad=True)
loss, _ = pcc_b200.chamfer_distance(x, torch.rand(2, 1200, 3, device=dev))
loss.backward()                                                     # grid Chamfer forward + backward kernels
torch.cuda.synchronize()
print("sanitize case done", float(loss), np.isfinite(x.grad.cpu().numpy()).all())
