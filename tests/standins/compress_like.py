"""Stand-in for the reference's compress.py + decompress.py call structure (see README.md): a script that parses argv and runs
at import, imports torchac / pytorch3d / pn_kit / AE at the top, and drives the model ONE PATCH AT A TIME through ae.sa and
ae.pn exactly like compress.py:113-122, then the decoder half like decompress.py:88-110.  Run through
`python -m pcc_b200.run tests/standins/compress_like.py ...`; writes every intermediate to --out (npz)."""
import argparse

import numpy as np
import torch
import torchac
from pytorch3d.loss import chamfer_distance
from pytorch3d.ops.knn import _KNN, knn_gather, knn_points  # noqa: F401

import AE
import pn_kit

torch.manual_seed(11)
np.random.seed(11)

parser = argparse.ArgumentParser(prog="compress_like.py")
parser.add_argument("cloud")
parser.add_argument("weights")
parser.add_argument("--out", required=True)
parser.add_argument("--N0", type=int, default=1024)
parser.add_argument("--ALPHA", type=int, default=2)
parser.add_argument("--K", type=int, default=256)
parser.add_argument("--d", type=int, default=16)
parser.add_argument("--L", type=int, default=7)
parser.add_argument("--device", default=torch.device("cuda" if torch.cuda.is_available() else "cpu"))
args = parser.parse_args()

K, k, B = args.K, args.K // args.ALPHA, 1
state = torch.load(args.weights, map_location=args.device)
ae = AE.AE(K=K, k=k, d=args.d, L=args.L).to(args.device)
ae.load_state_dict(state["ae"])
ae.eval()
prob = AE.ConditionalProbabilityModel(args.L, args.d).to(args.device)
prob.load_state_dict(state["prob"])
prob.eval()

with torch.no_grad():
    pc = torch.Tensor(np.load(args.cloud)).to(args.device).unsqueeze(0)
    pc, center, longest = pn_kit.normalize(pc, margin=0.01)
    N = pc.shape[1]
    S = int(N * args.ALPHA // K)
    fps_idx = pn_kit.farthest_point_sample_batch(pc, S)
    sampled_xyz = pn_kit.index_points(pc, fps_idx)
    octree_codes, sampled_bits = pn_kit.encode_sampled_np(sampled_xyz.detach().cpu().numpy(), scale=1, N=N,
                                                          min_bpp=pn_kit.OCTREE_BPP_DICT[K])
    rec_sampled_xyz = torch.Tensor(pn_kit.decode_sampled_np(octree_codes, scale=1)).to(args.device)
    assert rec_sampled_xyz.shape == sampled_xyz.shape
    dist, group_idx, grouped_xyz = knn_points(rec_sampled_xyz, pc, K=K, return_nn=True)
    grouped_xyz -= rec_sampled_xyz.view(B, S, 1, 3)
    scale = (N / args.N0) ** (1 / 3)
    x_patches = grouped_xyz.view(B * S, K, 3).transpose(1, 2) * scale
    feats = torch.cat([ae.sa(x_patches[j].view(1, 3, K))[1].cpu() for j in range(S)])
    latent = torch.cat([ae.pn(torch.cat((x_patches[j].unsqueeze(0), feats[j].to(args.device).unsqueeze(0)), dim=1)).cpu()
                        for j in range(S)])
    spread = ae.L - 0.2
    latent = torch.sigmoid(latent) * spread - spread / 2
    latent_quantized = ae.quantize(latent)
    pmf = prob(rec_sampled_xyz)
    cdf = pn_kit.pmf_to_cdf(pmf).cpu()
    sym = latent_quantized.view(B, S, -1).to(torch.int16).cpu() + args.L // 2
    byte_stream = torchac.encode_float_cdf(cdf, sym, check_input_bounds=True)
    # ---- the decoder side ----
    back = torchac.decode_float_cdf(cdf, byte_stream) - args.L // 2
    lq = back.view(S, -1).float().to(args.device)
    lin = ae.inv_pool(lq).view(B * S, -1, k)
    new_xyz = ae.inv_mlp(torch.cat((lin, lq.unsqueeze(-1).repeat((1, 1, k))), dim=1)).transpose(2, 1) / scale
    rec = (new_xyz.view(B, S, k, 3) + rec_sampled_xyz.view(B, S, 1, 3)).reshape(B, -1, 3)
    cham, _ = chamfer_distance(rec, pc)
    np.savez(args.out, pc=pc.cpu().numpy(), fps_idx=fps_idx.cpu().numpy(), centres=rec_sampled_xyz.cpu().numpy(), group_idx=group_idx.cpu().numpy(),
             feats=feats.float().numpy(), latent=latent.numpy(), latent_q=latent_quantized.numpy(), pmf=pmf.float().cpu().numpy(),
             nbytes=len(byte_stream), sym_back=back.numpy(), rec=rec.cpu().numpy(), chamfer=float(cham),
             sa_swapped=bool(getattr(type(ae.sa).forward, "__pcc_b200__", False)),
             prob_swapped=bool(getattr(type(prob).forward, "__pcc_b200__", False)))
print("done")
