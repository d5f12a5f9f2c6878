"""Stand-in for the reference's pn_kit.py (interface only; see README.md).  Imports the same third-party names at import
time as /root/reference/pn_kit.py:10 so that the pytorch3d shim is exercised."""
import torch
import torch.nn as nn
import torch.nn.functional as F
from pytorch3d.ops.knn import _KNN, knn_gather, knn_points  # noqa: F401


def _block(cin, cout, relu, bn):
    mods = [nn.Conv2d(cin, cout, 1)]
    if relu and bn:
        mods.append(nn.BatchNorm2d(cout))
    if relu:
        mods.append(nn.ReLU())
    return nn.Sequential(*mods)


class _Stack(nn.Module):
    def __init__(self, in_channel, mlps, relu, bn):
        super().__init__()
        dims = [in_channel] + list(mlps)
        self.mlp_Modules = nn.ModuleList(_block(dims[i], dims[i + 1], relu[i], bn) for i in range(len(dims) - 1))

    def _run(self, points):
        h = points.unsqueeze(-1)
        for m in self.mlp_Modules:
            h = m(h)
        return h.squeeze(-1)


class PointNet(_Stack):
    def forward(self, points):            # [B, C, N] -> [B, D]
        return self._run(points).max(dim=2)[0]


class MLP(_Stack):
    def forward(self, points):            # [B, C, N] -> [B, D, N]
        return self._run(points)


class SetAbstraction(nn.Module):
    def __init__(self, npoint, K, in_channel, mlp, bn=False, finalRelu=True):
        super().__init__()
        self.npoint, self.K, self.bn, self.finalRelu = npoint, K, bn, finalRelu
        if bn:
            self.bn0, self.bn1, self.bn2 = (nn.BatchNorm2d(c) for c in mlp)
        self.conv0 = nn.Conv2d(in_channel + 3, mlp[0], 1)
        self.conv1 = nn.Conv2d(mlp[0], mlp[1], 1)
        self.conv2 = nn.Conv2d(mlp[1], mlp[2], 1)

    def forward(self, xyz):               # [B, 3, N] -> ([B, 3, S], [B, D', S])
        p = xyz.permute(0, 2, 1)
        B, N, C = p.shape
        q = p if self.npoint == N else index_points(p, farthest_point_sample_batch(p, self.npoint))
        _, _, g = knn_points(q, p, K=self.K, return_nn=True)
        g = (g - q.view(B, self.npoint, 1, C)).permute(0, 3, 2, 1)
        for i in range(3):
            g = getattr(self, f"conv{i}")(g)
            if self.bn:
                g = getattr(self, f"bn{i}")(g)
            if i < 2 or self.finalRelu:
                g = F.relu(g)
        return q.permute(0, 2, 1), g.max(dim=2)[0]


def farthest_point_sample_batch(xyz, npoint):
    B, N, _ = xyz.shape
    out = torch.zeros(B, npoint, dtype=torch.long, device=xyz.device)
    dist = torch.full((B, N), 1e10, device=xyz.device)
    far = torch.randint(0, N, (B,), dtype=torch.long).to(xyz.device)
    rows = torch.arange(B, device=xyz.device)
    for i in range(npoint):
        out[:, i] = far
        d = ((xyz - xyz[rows, far].view(B, 1, 3)) ** 2).sum(-1)
        dist = torch.where(d < dist, d, dist)
        far = dist.max(-1)[1]
    return out


def index_points(points, idx):
    B = points.shape[0]
    rows = torch.arange(B, device=points.device).view([B] + [1] * (idx.dim() - 1)).expand_as(idx)
    return points[rows, idx, :]


OCTREE_BPP_DICT = {1024: 0.07, 512: 0.125, 256: 0.25, 128: 0.5, 64: 1.0}


def normalize(pc, margin=0.01):           # [1, N, 3]; bounding box of cloud 0
    lo, hi = pc[0].min(dim=0)[0], pc[0].max(dim=0)[0]
    center, longest = (hi + lo) / 2, (hi - lo).max()
    return (pc - center) * (1 - margin) / longest + 0.5, center, longest


def denormalize(pc, center, longest, margin=0.01):
    return (pc - 0.5) * longest / (1 - margin) + center


def pmf_to_cdf(pmf):
    cdf = pmf.cumsum(dim=-1)
    return torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1).clamp(max=1.)


def encode_sampled_np(sampled_xyz, scale, N, min_bpp):
    raise NotImplementedError("stand-in: the octree coder is provided by pcc_b200.install()")


def decode_sampled_np(codes, scale):
    raise NotImplementedError("stand-in: the octree coder is provided by pcc_b200.install()")
