"""Stand-in for the reference's PPPF_AE.py (interface only; see README.md)."""
import torch
import torch.nn as nn
import torch.nn.functional as F
from AE import STEQuantize
from pointnet_sa_module import PointnetSAModule
from pytorch3d.loss import chamfer_distance  # noqa: F401


class PointNetPP(nn.Module):
    def __init__(self, points=512, sa1_mlp=[64, 64, 128], sa2_mlp=[128, 128, 128, 256], sa3_mlp=[256, 256, 512],
                 feature_dim=1024, bn=False):
        super().__init__()
        self.sa1 = PointnetSAModule(npoint=points, radius=0.2, nsample=32, mlp=[3] + sa1_mlp, use_xyz=True, in_channels=0)
        self.sa2 = PointnetSAModule(npoint=128, radius=0.4, nsample=64, mlp=sa2_mlp, use_xyz=True, in_channels=128)
        self.sa3 = PointnetSAModule(npoint=32, radius=0.8, nsample=128, mlp=sa3_mlp + [feature_dim], use_xyz=True, in_channels=256)

    def forward(self, xyz, features=None):
        for sa in (self.sa1, self.sa2, self.sa3):
            xyz, features = sa(xyz, features)
        return xyz, features.max(dim=2)[0]


class FoldingNet(nn.Module):
    def __init__(self, points=512, grid_size=45, feature_dim=1024):
        super().__init__()
        self.grid_size, self.num_points, self.feature_dim, self.size = grid_size, grid_size * grid_size, feature_dim, points

        def stage(cin, width):
            return nn.Sequential(nn.Conv1d(cin, width, 1), nn.ReLU(), nn.Conv1d(width, width, 1), nn.ReLU(), nn.Conv1d(width, 3, 1))

        self.mlp1, self.mlp2 = stage(feature_dim + 2, points), stage(feature_dim + 3, 128)

    def build_grid(self, batch_points, device):
        t = torch.linspace(-1, 1, self.grid_size)
        g = torch.stack(torch.meshgrid(t, t, indexing="ij"), dim=-1).reshape(-1, 2)
        return g.unsqueeze(0).repeat(batch_points, 1, 1).to(device)

    def forward(self, latent_quantized):  # [B, F] -> [B, grid_size^2, 3]
        B = latent_quantized.size(0)
        tiled = latent_quantized.unsqueeze(1).repeat(1, self.num_points, 1)
        coarse = self.mlp1(torch.cat([self.build_grid(B, latent_quantized.device), tiled], dim=-1).transpose(2, 1))
        return self.mlp2(torch.cat([coarse, tiled.transpose(2, 1)], dim=1)).transpose(2, 1)


class PPPF_AE(nn.Module):
    def __init__(self, K=512, k=0, d=16, L=7, dim=1024):
        super().__init__()
        self.L = L
        self.encoder = PointNetPP(points=K, feature_dim=dim)
        self.decoder = FoldingNet(points=K, grid_size=d)
        self.enc_proj, self.dec_proj = nn.Linear(dim, d), nn.Linear(d, dim)
        self.quantize = STEQuantize.apply

    def forward(self, xyz):
        _, latent = self.encoder(xyz)
        spread = self.L - 0.2
        latent = torch.sigmoid(latent) * spread - spread / 2
        lq = self.quantize(self.enc_proj(latent))
        return self.decoder(self.dec_proj(lq)), latent, lq


class ConditionalProbabilityModel(nn.Module):
    def __init__(self, L, d):
        super().__init__()
        self.L, self.d = L, d
        self.model_pnpp = PointNetPP(sa1_mlp=[64, 64, 128], sa2_mlp=[128, 128, 256], sa3_mlp=[256, 512, 1024], bn=False)
        self.model_mlp = nn.Sequential(nn.Conv2d(3 + 1024, 512, 1), nn.ReLU(), nn.Conv2d(512, 512, 1), nn.ReLU(),
                                       nn.Conv2d(512, d * L, 1))

    def forward(self, sampled_xyz):
        B, S, _ = sampled_xyz.shape
        _, feature = self.model_pnpp(sampled_xyz)
        x = torch.cat((sampled_xyz, feature.unsqueeze(1).repeat(1, S, 1)), dim=2).unsqueeze(-1).transpose(1, 2)
        return F.softmax(self.model_mlp(x).transpose(1, 2).view(B, S, self.d, self.L), dim=3)


class AE(PPPF_AE):
    pass
