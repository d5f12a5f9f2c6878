"""CPU fp32 restatement of the reference's network bodies and of its compress / decompress / eval flow, on top of
the C oracle.  TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py cpu_baseline / --impl reference).

Every function takes the reference's state_dict (same keys) and follows the cited lines; the layouts and op
order are the reference's (channel-first F.conv2d on [B, C, K, S]), so the outputs match the reference modules
run on CPU with the oracle ops injected (tests/test_oracle.py::test_torch_modules_match_reference_ae_golden pins that against tests/golden/ae_modules.npz).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import oracle as orc


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def set_abstraction(sd, xyz, K=16, prefix="sa.", threads=1):
    """pn_kit.SetAbstraction.forward, /root/reference/pn_kit.py:164-211 with S == N (no FPS), bn=False.
    xyz [B, 3, N] -> features [B, D', N]."""
    p = xyz.permute(0, 2, 1).contiguous()
    B, N, C = p.shape
    _, _, g = orc.knn_points(p.numpy(), p.numpy(), K, True, threads=threads)     # :190
    g = _t(g) - p.view(B, N, 1, C)                                                # :191
    g = g.permute(0, 3, 2, 1)                                                     # :196  [B, 3, K, N]
    g = F.relu(F.conv2d(g, sd[prefix + "conv0.weight"], sd[prefix + "conv0.bias"]))  # :198
    g = F.relu(F.conv2d(g, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"]))  # :199
    g = F.relu(F.conv2d(g, sd[prefix + "conv2.weight"], sd[prefix + "conv2.bias"]))  # :201-205 (finalRelu)
    return torch.max(g, 2)[0]                                                     # :207


def _conv_chain(sd, x, prefix, n, relu):
    x = x.unsqueeze(-1)
    for i in range(n):
        x = F.conv2d(x, sd[f"{prefix}mlp_Modules.{i}.0.weight"], sd[f"{prefix}mlp_Modules.{i}.0.bias"])
        if relu[i]:
            x = F.relu(x)
    return x.squeeze(-1)


def pointnet(sd, points, prefix="pn."):
    """pn_kit.PointNet.forward, /root/reference/pn_kit.py:124-144.  points [B, C, N] -> [B, D]."""
    return torch.max(_conv_chain(sd, points, prefix, 4, [True, True, True, False]), 2)[0]


def ae_encode(sd, x_patches, L=7, threads=1):
    """Encoder half of AE.forward, /root/reference/AE.py:37-45.  x_patches [BS, K, 3] -> (latent, rounded)."""
    xyz = x_patches.transpose(2, 1)
    feat = set_abstraction(sd, xyz, threads=threads)
    latent = pointnet(sd, torch.cat((xyz, feat), dim=1))
    spread = L - 0.2
    latent = torch.sigmoid(latent) * spread - spread / 2
    return latent, latent.round()


def ae_decode(sd, latent_q, k=128):
    """Decoder half of AE.forward, /root/reference/AE.py:48-53.  latent_q [BS, d] -> [BS, k, 3]."""
    BS = latent_q.shape[0]
    x = latent_q
    for i in (0, 2, 4):
        x = F.relu(F.linear(x, sd[f"inv_pool.{i}.weight"], sd[f"inv_pool.{i}.bias"]))
    x = x.view(BS, -1, k)
    x = torch.cat((x, latent_q.unsqueeze(-1).repeat((1, 1, k))), dim=1)
    return _conv_chain(sd, x, "inv_mlp.", 4, [True, True, True, False]).transpose(2, 1)


def normalize(pc, margin=0.01):
    """pn_kit.normalize, /root/reference/pn_kit.py:47-60, one cloud [1, N, 3]."""
    x, y, z = pc[0, :, 0], pc[0, :, 1], pc[0, :, 2]
    center = torch.Tensor([(x.max() + x.min()) / 2, (y.max() + y.min()) / 2, (z.max() + z.min()) / 2])
    longest = torch.max(torch.Tensor([x.max() - x.min(), y.max() - y.min(), z.max() - z.min()]))
    pc = pc - center
    pc = pc * (1 - margin) / longest
    return pc + 0.5, center, longest


def quantise_centres(c, depth):
    """octree_np.getDecodeFromPc's quantisation, /root/reference/octree_np.py:114-133, FPS order kept."""
    cube = np.float32(1.0 / max(1.0, math.pow(2.0, min(depth, 30))))
    return (c // cube * cube) + (cube / 2)


OCTREE_BPP_DICT = {1024: 0.07, 512: 0.125, 256: 0.25, 128: 0.5, 64: 1.0}  # pn_kit.py:17-23


def compress_decompress_eval(sd, cloud, start_idx, K=256, k=128, d=16, L=7, N0=1024, alpha=2, centre_depth=6,
                             threads=1, per_patch_loop=False, centre_mode="fixed"):
    """One cloud [N,3] through the hot path of compress.py:90-127 -> decompress.py:96-116 -> eval.py:180,199-205.
    per_patch_loop=True feeds the encoder one patch at a time exactly as compress.py:113-122 does.
    centre_mode="coded" runs the octree centre coder of compress.py:98 (pn_kit.encode_sampled_np, C restatement pinned to
    the reference) and continues with the centres a correct decoder recovers from that stream."""
    pc = _t(cloud)[None]
    N = pc.shape[1]
    S = int(N * alpha // K)
    pcn, center, longest = normalize(pc)
    idx = orc.fps(pcn.numpy(), S, np.asarray([start_idx], dtype=np.int64), 1e10)
    octree = None
    if centre_mode == "coded":
        fps_xyz = orc.gather(pcn.numpy(), idx)
        codes, _, depths = orc.encode_sampled_np(fps_xyz, 1, N, OCTREE_BPP_DICT[K])
        octree = dict(bits=codes[0], depth=depths[0], bytes=orc.bits_to_bytes(codes[0]))
        centres = orc.octree_stream_centres(fps_xyz[0], depths[0], S)[None].astype(np.float32)
    else:
        centres = quantise_centres(orc.gather(pcn.numpy(), idx), centre_depth).astype(np.float32)
    _, _, nn = orc.knn_points(centres, pcn.numpy(), K, True, threads=threads)
    patches = (_t(nn) - _t(centres).view(1, S, 1, 3)).view(S, K, 3)
    scale = (N / N0) ** (1 / 3)
    patches = patches * scale
    with torch.no_grad():
        if per_patch_loop:
            outs = [ae_encode(sd, patches[j:j + 1], L, threads) for j in range(S)]
            latent = torch.cat([o[0] for o in outs])
            latent_q = torch.cat([o[1] for o in outs])
        else:
            latent, latent_q = ae_encode(sd, patches, L, threads)
        rec = ae_decode(sd, latent_q, k) / scale
    rec = (rec.view(1, S, -1, 3) + _t(centres).view(1, S, 1, 3)).reshape(1, -1, 3)
    rec = (rec - 0.5) * longest / (1 - 0.01) + center                            # pn_kit.denormalize :62-66
    rec_np, orig = rec[0].numpy(), np.ascontiguousarray(cloud)
    mn, mx = orig.min(), orig.max()                                              # eval.py:199-202
    cham = orc.chamfer(((rec_np - mn) / (mx - mn))[None], ((orig - mn) / (mx - mn))[None], threads=threads)[0]
    psnr, mse = orc.d1_psnr(orig, rec_np)                                        # eval.py:43-98 (float64)
    return dict(latent=latent.numpy(), latent_q=latent_q.numpy(), centres=centres[0], rec=rec_np, chamfer=cham,
                d1_psnr=psnr, d1_mse=mse, octree=octree)
