"""PointNet++ set abstraction and the PPPF auto-encoder on the B200 ops: same class names, constructor arguments and
state_dict keys as the reference (/root/reference/pointnet_sa_module.py:37-93, /root/reference/PPPF_AE.py:9-150), so its
checkpoints load unchanged; the forward bodies are batched device code:

    FPS (start 0, fused centre gather) -> ball query (first nsample in radius, -1 -> point 0) -> row gather ->
    Conv2d+BatchNorm(eval, folded)+ReLU stack on the tensor cores -> max over nsample   (one fused chain per run of
    layers whose weights fit in shared memory; wider layers are library GEMMs, see mlp_ops.run_chain).

With autograd enabled and trainable parameters (training), every module runs a differentiable body instead: FPS, ball query
and the row gathers stay on the pcc kernels (no gradient flows into the indices; the gather has a hand-written backward), the
Conv2d + BatchNorm (batch statistics) + ReLU stacks and FoldingNet's Conv1d layers run as torch layers under autograd.
"""


def _training_pass(module):
    return torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters())

import torch
import torch.nn as nn

from . import mlp_ops, ops
from .modules import STEQuantize


def _fold_bn(conv, bn):
    """Eval-mode BatchNorm folded into the preceding 1x1 convolution: y = (Wx + b - mean) * gamma / sqrt(var + eps) + beta."""
    w = conv.weight.flatten(1)
    b = conv.bias if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    return (w * scale[:, None]).contiguous(), ((b - bn.running_mean) * scale + bn.bias).contiguous()


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
        self._folded = None

    def folded_layers(self):
        key = tuple(p._version for p in self.parameters()) + tuple(b._version for b in self.buffers())
        if self._folded is None or self._folded[0] != key:
            mods = list(self.mlp)
            self._folded = (key, [_fold_bn(mods[i], mods[i + 1]) + (True,) for i in range(0, len(mods), 3)])
        return self._folded[1]

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
            return self._forward_fused(xyz, features)

    def _forward_fused(self, xyz, features=None):
        B, N, _ = xyz.shape
        fps_idx, new_xyz = ops.fps(xyz, self.npoint, None, ops.FLT_MAX, return_xyz=True)   # :66-68 (start index 0)
        if self.npoint > N:                                                                  # :67 clamp: pads read point 0
            new_xyz = ops.gather(xyz, fps_idx.clamp(min=0))
        _, idx = ops.ball_query(new_xyz, xyz, self.nsample, self.radius, return_dists=False)  # :71
        idx = idx.clamp(min=0)                                                               # :27 (pads -> point 0)
        layers = self.folded_layers()
        rows = B * self.npoint * self.nsample
        cin = (features.shape[1] if features is not None else 0) + (3 if self.use_xyz else 0)
        if cin >= 64 and all(mlp_ops.linear_supported(rows, w.shape[0]) for w, _, _ in layers) and \
                mlp_ops.linear_supported(rows, layers[-1][0].shape[0], self.nsample):
            # wide stack: one grouping pass (gather + cat + bf16, zero padded to the GEMM's K granule), then every layer on
            # the streamed tensor-core GEMM with the max over nsample fused into the last one       :73-91
            a = mlp_ops.gather_concat_bf16(features.permute(0, 2, 1) if features is not None else None,
                                           xyz if self.use_xyz else None, idx, (cin + 63) // 64 * 64)
            out = mlp_ops.stream_chain(a, layers, group=self.nsample)
            return new_xyz, out.view(B, self.npoint, -1).permute(0, 2, 1)
        segs = []
        if features is not None:
            segs.append((ops.gather(features.permute(0, 2, 1).contiguous(), idx).view(-1, features.shape[1]), 1))  # :74-77
        if self.use_xyz:
            segs.append((ops.gather(xyz, idx).view(-1, 3), 1))                                # :80-85 (not recentred)
        out = mlp_ops.run_chain(segs, layers, group=self.nsample)                           # :89-91
        return new_xyz, out.view(B, self.npoint, -1).permute(0, 2, 1)


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
    """PPPF_AE.FoldingNet (/root/reference/PPPF_AE.py:50-109).  The latent is the same for every grid point, so the
    first layer of each folding stage is a per-cloud vector (latent part, a small library GEMM) plus a 2- or 3-wide
    per-point term; the remaining Conv1d layers are plain [B*N, K] x [K, K] library GEMMs in bf16."""

    def __init__(self, points=512, grid_size=45, feature_dim=1024):
        super().__init__()
        self.grid_size, self.num_points, self.feature_dim, self.size = grid_size, grid_size * grid_size, feature_dim, points
        self.mlp1 = nn.Sequential(nn.Conv1d(feature_dim + 2, points, 1), nn.ReLU(), nn.Conv1d(points, points, 1), nn.ReLU(),
                                  nn.Conv1d(points, 3, 1))
        self.mlp2 = nn.Sequential(nn.Conv1d(feature_dim + 3, 128, 1), nn.ReLU(), nn.Conv1d(128, 128, 1), nn.ReLU(),
                                  nn.Conv1d(128, 3, 1))

    def build_grid(self, batch_points, device):
        """PPPF_AE.py:80-89; the grid is a constant: built once per (batch, device) and kept on the device."""
        key = (batch_points, str(device))
        cached = getattr(self, "_grid_cache", None)
        if cached is None or cached[0] != key:
            x = torch.linspace(-1, 1, self.grid_size)
            gx, gy = torch.meshgrid(x, x, indexing="ij")
            grid = torch.stack([gx, gy], dim=-1).reshape(-1, 2).unsqueeze(0).repeat(batch_points, 1, 1).to(device)
            cached = self._grid_cache = (key, grid)
        return cached[1]

    @staticmethod
    def _stage(mlp, local, latent, n_local, latent_first):
        """One folding stage on [B, N, n_local] per-point inputs and a [B, F] latent."""
        w0 = mlp[0].weight.squeeze(-1)
        if latent_first:   # mlp2: cat([coarse(3), latent]) -> columns [0:n_local] local, rest latent
            w_loc, w_lat = w0[:, :n_local], w0[:, n_local:]
        else:              # mlp1: cat([grid(2), latent])
            w_loc, w_lat = w0[:, :n_local], w0[:, n_local:]
        B, N, _ = local.shape
        per_cloud = torch.addmm(mlp[0].bias, latent, w_lat.t())                      # [B, K]
        h = torch.relu(local @ w_loc.t() + per_cloud[:, None, :]).reshape(B * N, -1)
        h = mlp_ops.run_chain(h, [(mlp[2].weight.squeeze(-1), mlp[2].bias, True), (mlp[4].weight.squeeze(-1), mlp[4].bias, False)])
        return h.view(B, N, 3)

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
            B = latent_quantized.size(0)
            grid = self.build_grid(B, latent_quantized.device)                           # [B, N, 2]
            coarse = self._stage(self.mlp1, grid, latent_quantized, 2, False)            # PPPF_AE.py:100-104
            return self._stage(self.mlp2, coarse, latent_quantized, 3, True)             # PPPF_AE.py:106-109


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
        with torch.set_grad_enabled(_training_pass(self)):
            return self._forward(xyz)

    def _forward(self, xyz):
        _, latent = self.encoder(xyz)
        spread = self.L - 0.2
        latent = torch.sigmoid(latent) * spread - spread / 2
        latent_quantized = self.quantize(self.enc_proj(latent))
        recon = self.decoder(self.dec_proj(latent_quantized))
        return recon, latent, latent_quantized


AE = PPPF_AE  # PPPF_AE.py:230-232
