"""PointNet++ set abstraction and the PPPF auto-encoder on the B200 ops: same class names, constructor arguments and
state_dict keys as the reference (/root/reference/pointnet_sa_module.py:37-93, /root/reference/PPPF_AE.py:9-150), so its
checkpoints load unchanged; the forward bodies are batched device code (bodies.py):

    FPS (start 0, fused centre gather) -> ball query (first nsample in radius, -1 -> point 0) -> row gather ->
    Conv2d+BatchNorm(eval, folded)+ReLU stack on the tensor cores -> max over nsample   (one fused chain per run of
    layers whose weights fit in shared memory; wider layers on the streamed tcgen05 GEMM, see mlp_ops.run_chain);
    FoldingNet: per-cloud latent term (fp32) + per-point term -> streamed GEMM -> fused tail.  No library GEMM.

With autograd enabled and trainable parameters (training), every module runs a differentiable body instead: FPS, ball query
and the row gathers stay on the pcc kernels (no gradient flows into the indices; the gather has a hand-written backward), the
Conv2d + BatchNorm (batch statistics) + ReLU stacks and FoldingNet's Conv1d layers run as torch layers under autograd.
"""
import torch
import torch.nn as nn

from . import bodies, ops
from .bodies import training_pass as _training_pass
from .modules import STEQuantize


class PointnetSAModule(nn.Module):
    """pointnet_sa_module.PointnetSAModule(npoint, radius, nsample, mlp, use_xyz=True, in_channels=0)."""

    def __init__(self, npoint, radius, nsample, mlp, use_xyz=True, in_channels=0):
        super().__init__()
        self.npoint, self.radius, self.nsample, self.use_xyz = npoint, radius, nsample, use_xyz
        last = in_channels + (3 if use_xyz else 0)
        layers = []
        for out_channel in mlp:
            layers += [nn.Conv2d(last, out_channel, 1), nn.BatchNorm2d(out_channel), nn.ReLU(inplace=True)]
            last = out_channel
        self.mlp = nn.Sequential(*layers)

    def forward_train(self, xyz, features=None):
        """pointnet_sa_module.py:58-93 under autograd: sampling / ball query / grouping on the pcc kernels, the shared MLP
        (Conv2d + BatchNorm with batch statistics when self.training + ReLU) as torch layers, max over nsample."""
        B, N, _ = xyz.shape
        with torch.no_grad():
            fps_idx, _ = ops.fps(xyz, self.npoint, None, ops.FLT_MAX, return_xyz=True)       # :66
            fps_idx = fps_idx.clamp(min=0)                                                  # :67
        new_xyz = ops.gather(xyz, fps_idx)                                                  # :68
        with torch.no_grad():
            _, idx = ops.ball_query(new_xyz, xyz, self.nsample, self.radius, return_dists=False)  # :71
            idx = idx.clamp(min=0)                                                          # :27
        parts = []
        if features is not None:
            parts.append(ops.gather(features.permute(0, 2, 1).contiguous(), idx))          # :74-77  [B, npoint, nsample, C]
        if self.use_xyz:
            parts.append(ops.gather(xyz, idx))                                              # :80-85
        grouped = torch.cat(parts, dim=-1).permute(0, 3, 1, 2)                              # :88  [B, C, npoint, nsample]
        return new_xyz, torch.max(self.mlp(grouped), 3)[0]                                  # :89-91

    def forward(self, xyz, features=None):
        """xyz [B,N,3], features [B,C,N] or None -> (new_xyz [B,npoint,3], new_features [B,C_out,npoint])."""
        if _training_pass(self):
            return self.forward_train(xyz, features)
        if self.training:
            raise NotImplementedError("pcc_b200.PointnetSAModule: train-mode BatchNorm without autograd is not built "
                                      "(call .eval() for inference)")
        with torch.no_grad():
            return bodies.sa_module_forward(self, xyz, features)


class PointNetPP(nn.Module):
    """PPPF_AE.PointNetPP (/root/reference/PPPF_AE.py:9-46)."""

    def __init__(self, points=512, sa1_mlp=(64, 64, 128), sa2_mlp=(128, 128, 128, 256), sa3_mlp=(256, 256, 512),
                 feature_dim=1024, bn=False):
        super().__init__()
        self.sa1 = PointnetSAModule(points, 0.2, 32, [3] + list(sa1_mlp), True, 0)          # PPPF_AE.py:29-31 (extra 3->3)
        self.sa2 = PointnetSAModule(128, 0.4, 64, list(sa2_mlp), True, 128)
        self.sa3 = PointnetSAModule(32, 0.8, 128, list(sa3_mlp) + [feature_dim], True, 256)

    def forward(self, xyz, features=None):
        xyz, features = self.sa1(xyz, features)
        xyz, features = self.sa2(xyz, features)
        xyz, features = self.sa3(xyz, features)
        return xyz, torch.max(features, dim=2)[0]


class FoldingNet(nn.Module):
    """PPPF_AE.FoldingNet (/root/reference/PPPF_AE.py:50-109).  The latent is the same for every grid point, so the first
    layer of each folding stage is a per-cloud vector (pcc_linear_small_f32) plus a 2- or 3-wide per-point term
    (pcc_fold_first_bf16); the 512 -> 512 layer runs on the streamed tcgen05 GEMM and the narrow tail on the fused chain."""

    def __init__(self, points=512, grid_size=45, feature_dim=1024):
        super().__init__()
        self.grid_size, self.num_points, self.feature_dim, self.size = grid_size, grid_size * grid_size, feature_dim, points
        self.mlp1 = nn.Sequential(nn.Conv1d(feature_dim + 2, points, 1), nn.ReLU(), nn.Conv1d(points, points, 1), nn.ReLU(),
                                  nn.Conv1d(points, 3, 1))
        self.mlp2 = nn.Sequential(nn.Conv1d(feature_dim + 3, 128, 1), nn.ReLU(), nn.Conv1d(128, 128, 1), nn.ReLU(),
                                  nn.Conv1d(128, 3, 1))

    def build_grid(self, batch_points, device):
        """PPPF_AE.py:80-89; the grid is a constant: built once per (batch, device) and kept on the device."""
        return bodies._folding_grid(self, batch_points, device)

    def forward_train(self, latent_quantized):
        """PPPF_AE.py:91-109 under autograd (torch Conv1d layers)."""
        B = latent_quantized.size(0)
        grid = self.build_grid(B, latent_quantized.device)
        latent_expanded = latent_quantized.unsqueeze(1).repeat(1, self.num_points, 1)
        coarse = self.mlp1(torch.cat([grid, latent_expanded], dim=-1).transpose(2, 1))
        fine = self.mlp2(torch.cat([coarse, latent_expanded.transpose(2, 1)], dim=1))
        return fine.transpose(2, 1)

    def forward(self, latent_quantized):
        if _training_pass(self):
            return self.forward_train(latent_quantized)
        with torch.no_grad():
            return bodies.folding_forward(self, latent_quantized)


class PPPF_AE(nn.Module):
    """PPPF_AE.PPPF_AE(K, k, d, L, dim) (/root/reference/PPPF_AE.py:114-150)."""

    def __init__(self, K=512, k=0, d=16, L=7, dim=1024):
        super().__init__()
        self.L = L
        self.encoder = PointNetPP(points=K, feature_dim=dim)
        self.decoder = FoldingNet(points=K, grid_size=d)
        self.enc_proj = nn.Linear(dim, d)
        self.dec_proj = nn.Linear(d, dim)
        self.quantize = STEQuantize.apply

    def forward(self, xyz):
        """PPPF_AE.py:131-150; differentiable when autograd is on and the parameters are trainable, fused otherwise."""
        if _training_pass(self):
            return self._forward_train(xyz)
        if bodies.has_train_mode_bn(self):
            raise NotImplementedError("pcc_b200.PPPF_AE: train-mode BatchNorm without autograd is not built (call .eval())")
        with torch.no_grad():
            return bodies.pppf_forward(self, xyz)

    def _forward_train(self, xyz):
        _, latent = self.encoder(xyz)
        spread = self.L - 0.2
        latent = torch.sigmoid(latent) * spread - spread / 2
        latent_quantized = self.quantize(self.enc_proj(latent))
        recon = self.decoder(self.dec_proj(latent_quantized))
        return recon, latent, latent_quantized


class ConditionalProbabilityModel(nn.Module):
    """PPPF_AE.ConditionalProbabilityModel(L, d) (/root/reference/PPPF_AE.py:181-228): PointNet++ backbone over the sampled
    points, then Conv2d 3+1024 -> 512 -> 512 -> d*L and a softmax over L."""

    def __init__(self, L, d):
        super().__init__()
        self.L, self.d = L, d
        self.model_pnpp = PointNetPP(sa1_mlp=[64, 64, 128], sa2_mlp=[128, 128, 256], sa3_mlp=[256, 512, 1024], bn=False)
        self.model_mlp = nn.Sequential(nn.Conv2d(3 + 1024, 512, 1), nn.ReLU(), nn.Conv2d(512, 512, 1), nn.ReLU(),
                                       nn.Conv2d(512, d * L, 1))

    def forward_train(self, sampled_xyz):
        B, S, _ = sampled_xyz.shape
        _, feature = self.model_pnpp(sampled_xyz)                                       # :207
        x = torch.cat((sampled_xyz, feature.unsqueeze(1).repeat(1, S, 1)), dim=2)       # :210-213
        out = self.model_mlp(x.unsqueeze(-1).transpose(1, 2))                           # :216-219
        return torch.softmax(out.transpose(1, 2).reshape(B, S, self.d, self.L), dim=3)  # :222-223

    def forward(self, sampled_xyz):
        if _training_pass(self):
            return self.forward_train(sampled_xyz)
        if bodies.has_train_mode_bn(self):
            raise NotImplementedError("pcc_b200.pppf.ConditionalProbabilityModel: call .eval() for inference")
        with torch.no_grad():
            return bodies.pppf_prob_forward(self, sampled_xyz)


AE = PPPF_AE  # PPPF_AE.py:230-232
