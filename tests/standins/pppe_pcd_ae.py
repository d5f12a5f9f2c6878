"""Stand-in for the live half of the reference's pppe_pcd_ae.py (interface only; see README.md): the PointNet++ encoder classes
PointNetSetAbstraction / PointNetSetAbstractionMSG / PointNet2EncoderFull with the reference's constructor arguments, attribute
names, state_dict keys and forward layouts; compact eager-torch bodies written for this repository."""
import torch
import torch.nn as nn
from pytorch3d.loss import chamfer_distance  # noqa: F401
from pytorch3d.ops.knn import knn_points

from pn_kit import farthest_point_sample_batch, index_points


def _block(cin, cout, bn, dims):
    conv, norm = (nn.Conv2d, nn.BatchNorm2d) if dims == 2 else (nn.Conv1d, nn.BatchNorm1d)
    if bn:
        return nn.Sequential(conv(cin, cout, kernel_size=1, bias=False), norm(cout), nn.ReLU(inplace=True))
    return nn.Sequential(conv(cin, cout, 1), nn.ReLU())


class PointNetSetAbstraction(nn.Module):
    def __init__(self, npoint, K, in_channel, mlp, bn=True):
        super().__init__()
        self.npoint, self.K = npoint, K
        chans = [in_channel + 3] + list(mlp)
        self.mlp_stack = nn.ModuleList(_block(a, b, bn, 2) for a, b in zip(chans[:-1], chans[1:]))

    def forward(self, xyz, points=None):   # xyz [B,N,3], points [B,C,N] -> ([B,S,3], [B,C_out,S])
        B, N, _ = xyz.shape
        centres = xyz if self.npoint == N else index_points(xyz, farthest_point_sample_batch(xyz, self.npoint))
        _, nbr, near = knn_points(centres.float(), xyz.float(), K=self.K, return_nn=True)
        g = near - centres[:, :, None, :]
        if points is not None:
            g = torch.cat((g, index_points(points.transpose(1, 2), nbr)), dim=-1)
        g = g.permute(0, 3, 2, 1).contiguous()
        for layer in self.mlp_stack:
            g = layer(g)
        return centres, g.max(dim=2)[0]


class PointNetSetAbstractionMSG(nn.Module):
    def __init__(self, npoint, scales, in_channel, bn=True):
        super().__init__()
        self.branches = nn.ModuleList(PointNetSetAbstraction(npoint, s["K"], in_channel, s["mlp"], bn=bn) for s in scales)

    def forward(self, xyz, points=None):
        res = [b(xyz, points) for b in self.branches]
        return res[-1][0], torch.cat([f for _, f in res], dim=1)


class PointNet2EncoderFull(nn.Module):
    def __init__(self, sa_blocks=None, latent_dim=256, bn=True):
        super().__init__()
        if sa_blocks is None:
            sa_blocks = [{"type": "MSG", "npoint": 512, "in_channel": 0,
                          "scales": [{"K": 16, "mlp": [32, 32, 64]}, {"K": 32, "mlp": [64, 64, 128]}]},
                         {"type": "SS", "npoint": 128, "K": 32, "mlp": [128, 128, 256], "in_channel": 192},
                         {"type": "SS", "npoint": 32, "K": 32, "mlp": [256, 256, 512], "in_channel": 256}]
        self.sa_modules = nn.ModuleList(
            PointNetSetAbstractionMSG(b["npoint"], b["scales"], b.get("in_channel", 0), bn=bn) if b["type"] == "MSG" else
            PointNetSetAbstraction(b["npoint"], b["K"], b.get("in_channel", 0), b["mlp"], bn=bn) for b in sa_blocks)
        last = sa_blocks[-1]
        width = sum(s["mlp"][-1] for s in last["scales"]) if last["type"] == "MSG" else last["mlp"][-1]
        head = _block(width, width, True, 1)
        self.global_conv = nn.Sequential(head[0], head[1], head[2], nn.Conv1d(width, latent_dim, 1))
        self.latent_dim = latent_dim

    def forward(self, x):                  # x [B,N,3] -> (latent [B,latent_dim], pooled [B,C_out])
        xyz, pts = x, None
        for sa in self.sa_modules:
            xyz, pts = sa(xyz, pts)
        pooled = pts.max(dim=2)[0]
        return self.global_conv(pooled[:, :, None])[:, :, 0], pooled
