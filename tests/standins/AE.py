"""Stand-in for the reference's AE.py (interface only; see README.md)."""
import torch
import torch.nn as nn
import torch.nn.functional as F
from pytorch3d.loss import chamfer_distance  # noqa: F401
from pn_kit import MLP, PointNet, SetAbstraction


class STEQuantize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.round()

    @staticmethod
    def backward(ctx, g):
        return g


class AE(nn.Module):
    def __init__(self, K, k, d, L):
        super().__init__()
        self.sa = SetAbstraction(npoint=K, K=16, in_channel=0, mlp=[32, 64, 128], bn=False)
        self.pn = PointNet(in_channel=3 + 128, mlps=[128, 256, 512, d], relu=[True, True, True, False], bn=False)
        self.inv_pool = nn.Sequential(nn.Linear(d, 256), nn.ReLU(), nn.Linear(256, 1024), nn.ReLU(), nn.Linear(1024, k * 128),
                                      nn.ReLU())
        self.inv_mlp = MLP(in_channel=d + 128, mlps=[128, 64, 32, 3], relu=[True, True, True, False], bn=False)
        self.K, self.k, self.L = K, k, L
        self.quantize = STEQuantize.apply

    def forward(self, xyz):               # [BS, K, 3] -> ([BS, k, 3], latent, latent_q)
        BS = xyz.shape[0]
        cf = xyz.transpose(2, 1)
        _, feat = self.sa(cf)
        latent = self.pn(torch.cat((cf, feat), dim=1))
        spread = self.L - 0.2
        latent = torch.sigmoid(latent) * spread - spread / 2
        lq = self.quantize(latent)
        lin = self.inv_pool(lq).view(BS, -1, self.k)
        out = self.inv_mlp(torch.cat((lin, lq.unsqueeze(-1).repeat((1, 1, self.k))), dim=1))
        return out.transpose(2, 1), latent, lq


class ConditionalProbabilityModel(nn.Module):
    def __init__(self, L, d):
        super().__init__()
        self.L, self.d = L, d
        self.model_pn = PointNet(in_channel=3, mlps=[64, 128, 256], relu=[True, True, True], bn=False)
        self.model_mlp = nn.Sequential(nn.Conv2d(3 + 256, 512, 1), nn.ReLU(), nn.Conv2d(512, 512, 1), nn.ReLU(),
                                       nn.Conv2d(512, d * L, 1))

    def forward(self, sampled_xyz):       # [B, S, 3] -> pmf [B, S, d, L]
        B, S, _ = sampled_xyz.shape
        feature = self.model_pn(sampled_xyz.transpose(1, 2))
        x = torch.cat((sampled_xyz, feature.repeat((1, S)).view(B, S, -1)), dim=2).unsqueeze(-1).transpose(1, 2)
        return F.softmax(self.model_mlp(x).transpose(1, 2).view(B, S, self.d, self.L), dim=3)
