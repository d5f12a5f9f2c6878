"""Stand-in for the reference's pointnet_sa_module.py (interface only; see README.md)."""
import torch
import torch.nn as nn
from pytorch3d.ops import ball_query, knn_gather, knn_points, sample_farthest_points


class PointnetPPOps:
    @staticmethod
    def furthest_point_sample(xyz, npoint):
        return sample_farthest_points(xyz, K=npoint)[1]

    @staticmethod
    def ball_query(radius, nsample, xyz, new_xyz):
        return ball_query(new_xyz, xyz, K=nsample, radius=radius)

    @staticmethod
    def group_points(features, idx):
        idx = getattr(idx, "idx", idx)
        return knn_gather(features, idx.clamp(min=0))

    @staticmethod
    def knn_point(k, xyz, new_xyz):
        r = knn_points(new_xyz, xyz, K=k, return_nn=True)
        return r[0], r[1]


class PointnetSAModule(nn.Module):
    def __init__(self, npoint, radius, nsample, mlp, use_xyz=True, in_channels=0):
        super().__init__()
        self.npoint, self.radius, self.nsample, self.use_xyz = npoint, radius, nsample, use_xyz
        last, layers = in_channels + (3 if use_xyz else 0), []
        for c in mlp:
            layers += [nn.Conv2d(last, c, 1), nn.BatchNorm2d(c), nn.ReLU(inplace=True)]
            last = c
        self.mlp = nn.Sequential(*layers)

    def forward(self, xyz, features=None):   # xyz [B,N,3], features [B,C,N] -> ([B,npoint,3], [B,C_out,npoint])
        idx = PointnetPPOps.furthest_point_sample(xyz, self.npoint).clamp(min=0)
        new_xyz = torch.gather(xyz, 1, idx.unsqueeze(-1).expand(-1, -1, 3))
        nbr = PointnetPPOps.ball_query(self.radius, self.nsample, xyz, new_xyz)
        parts = []
        if features is not None:
            parts.append(PointnetPPOps.group_points(features.permute(0, 2, 1), nbr))
        if self.use_xyz:
            parts.append(PointnetPPOps.group_points(xyz, nbr))
        g = torch.cat(parts, dim=-1).permute(0, 3, 1, 2)
        return new_xyz, self.mlp(g).max(dim=3)[0]
