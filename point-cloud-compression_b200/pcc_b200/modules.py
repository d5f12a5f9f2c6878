"""Parameter containers with the reference's state_dict keys, and the batched device forward of the IPDAE patch
auto-encoder (/root/reference/AE.py:12-55) built on the pcc_b200 kernels.

The containers mirror the reference classes only as far as names, constructor arguments and parameter keys /
shapes go, so `ae.pkl` checkpoints written by the reference's train.py load unchanged:
    sa.conv{0,1,2}.{weight,bias}                 pn_kit.SetAbstraction   (pn_kit.py:146-162)
    pn.mlp_Modules.{0..3}.0.{weight,bias}        pn_kit.PointNet         (pn_kit.py:98-121)
    inv_pool.{0,2,4}.{weight,bias}               AE.inv_pool             (AE.py:19-26)
    inv_mlp.mlp_Modules.{0..3}.0.{weight,bias}   pn_kit.MLP              (pn_kit.py:263-287)
The forward bodies are NOT the reference's: they run channel-last on flattened [rows, C] activations and call the
fused kernels (in-patch kNN -> shared MLP -> max over neighbours, ...).
"""
import torch
import torch.nn as nn

from . import mlp_ops, ops


class STEQuantize(torch.autograd.Function):
    """Round in the forward pass, identity gradient (AE.STEQuantize, /root/reference/AE.py:72-85)."""

    @staticmethod
    def forward(ctx, x):
        return x.round()

    @staticmethod
    def backward(ctx, grad_outputs):
        return grad_outputs


def _conv_stack(channels):
    mods = nn.ModuleList()
    for cin, cout in zip(channels[:-1], channels[1:]):
        mods.append(nn.Sequential(nn.Conv2d(cin, cout, 1)))  # key "<i>.0.weight", ReLU has no parameters
    return mods


class SetAbstraction(nn.Module):
    """pn_kit.SetAbstraction(npoint, K, in_channel, mlp, bn=False, finalRelu=True) with S == N (no FPS)."""

    def __init__(self, npoint, K, in_channel, mlp, bn=False, finalRelu=True):
        super().__init__()
        if bn:
            raise NotImplementedError("pcc_b200.SetAbstraction: bn=True is not used by the reference AE")
        self.npoint, self.K, self.finalRelu = npoint, K, finalRelu
        self.conv0 = nn.Conv2d(in_channel + 3, mlp[0], 1)
        self.conv1 = nn.Conv2d(mlp[0], mlp[1], 1)
        self.conv2 = nn.Conv2d(mlp[1], mlp[2], 1)

    def layers(self):
        return [(self.conv0.weight.flatten(1), self.conv0.bias, True), (self.conv1.weight.flatten(1), self.conv1.bias, True),
                (self.conv2.weight.flatten(1), self.conv2.bias, self.finalRelu)]

    def forward_points(self, xyz, out_dtype=torch.float32):
        """xyz [BS, P, 3] (channel-last) -> per-point features [BS, P, C_out] (channel-last).
        pn_kit.py:164-211: kNN(K) in the patch, recentre on the query, shared MLP, max over the K neighbours."""
        BS, P, _ = xyz.shape
        if self.npoint != P:
            raise NotImplementedError("pcc_b200.SetAbstraction: only the S == N configuration of AE.py:16 is built")
        _, _, grouped = ops.knn(xyz, xyz, self.K, return_nn=True, centre_sub=True, nn_only=True)  # [BS,P,K,3]
        return mlp_ops.fused_chain(grouped.reshape(BS * P * self.K, 3), self.layers(), group=self.K,
                                   out_dtype=out_dtype).reshape(BS, P, -1)

    def forward_points_train(self, xyz):
        """Differentiable body (training): the kNN grouping runs on the pcc kernel (no gradient flows into it, SURVEY.md
        3.3), the shared MLP is issued as fp32 library GEMMs under autograd, like the reference's default fp32 training."""
        BS, P, _ = xyz.shape
        _, _, grouped = ops.knn(xyz.detach(), xyz.detach(), self.K, return_nn=True, centre_sub=True, nn_only=True)
        h = grouped.reshape(BS * P * self.K, 3)
        for w, b, relu in self.layers():
            h = torch.addmm(b, h, w.t())
            if relu:
                h = torch.relu(h)
        return h.view(BS * P, self.K, -1).max(dim=1)[0].view(BS, P, -1)

    def forward(self, xyz):
        """Reference signature: xyz [B, 3, N] -> (new_xyz [B, 3, S], new_points [B, D', S])."""
        feat = self.forward_points(xyz.permute(0, 2, 1).contiguous())
        return xyz, feat.permute(0, 2, 1)


class PointNet(nn.Module):
    """pn_kit.PointNet(in_channel, mlps, relu, bn): shared MLP then max over the points."""

    def __init__(self, in_channel, mlps, relu, bn=False):
        super().__init__()
        if bn:
            raise NotImplementedError("pcc_b200.PointNet: bn=True is not used by the reference AE")
        self.relu = list(relu)
        self.mlp_Modules = _conv_stack([in_channel] + list(mlps))

    def layers(self):
        return [(m[0].weight.flatten(1), m[0].bias, r) for m, r in zip(self.mlp_Modules, self.relu)]

    def forward_points(self, x):
        """x [BS, P, C] channel-last -> [BS, D]."""
        BS, P, C = x.shape
        return mlp_ops.mlp_chain_groupmax(x.reshape(BS * P, C), self.layers(), group=P)

    def forward_xyz_feat(self, xyz, feat):
        """The AE.py:39 call `pn(cat((xyz, feat)))` without materialising the concatenation: xyz [BS,P,3] fp32 and
        feat [BS,P,F] (bf16 from the SetAbstraction kernel) are two input segments of the fused chain; the first
        layer's weight columns are rotated once so the 16-byte aligned feature block comes first."""
        BS, P, F = feat.shape
        layers = self.layers()
        w0 = layers[0][0]
        key = (w0.data_ptr(), w0._version)
        if getattr(self, "_rot_key", None) != key:
            self._rot_w0 = torch.cat((w0[:, 3:], w0[:, :3]), dim=1).detach().contiguous()
            self._rot_key = key
        layers[0] = (self._rot_w0, layers[0][1], layers[0][2])
        n = mlp_ops._split(layers)
        if n == 0:
            raise RuntimeError("pcc_b200.PointNet: first layer does not fit the fused kernel")
        h = mlp_ops.fused_chain([(feat.reshape(BS * P, F), 1), (xyz.reshape(BS * P, 3), 1)], layers[:n],
                                group=P if n == len(layers) else 0,
                                out_dtype=torch.float32 if n == len(layers) else torch.bfloat16)
        if n == len(layers):
            return h
        if mlp_ops.pn_tail_supported(h, layers[n:], P):   # 256 -> 512 -> d + max over the patch, one launch
            return mlp_ops.pn_tail(h, layers[n:])
        h = mlp_ops.library_chain(h, layers[n:])
        return h.view(BS, P, -1).max(dim=1)[0]

    def forward(self, points):
        """Reference signature: points [B, C, N] -> [B, D]."""
        return self.forward_points(points.permute(0, 2, 1).contiguous())


class MLP(nn.Module):
    """pn_kit.MLP(in_channel, mlps, relu, bn): shared MLP, no pooling."""

    def __init__(self, in_channel, mlps, relu, bn=False):
        super().__init__()
        if bn:
            raise NotImplementedError("pcc_b200.MLP: bn=True is not used by the reference AE")
        self.relu = list(relu)
        self.mlp_Modules = _conv_stack([in_channel] + list(mlps))

    def layers(self):
        return [(m[0].weight.flatten(1), m[0].bias, r) for m, r in zip(self.mlp_Modules, self.relu)]

    def forward_points(self, x):
        BS, P, C = x.shape
        return mlp_ops.mlp_chain(x.reshape(BS * P, C), self.layers()).reshape(BS, P, -1)

    def forward(self, points):
        return self.forward_points(points.permute(0, 2, 1).contiguous()).permute(0, 2, 1)


class AE(nn.Module):
    """AE.AE(K, k, d, L) (/root/reference/AE.py:12-55): IPDAE patch auto-encoder, same parameters, batched forward."""

    def __init__(self, K, k, d, L):
        super().__init__()
        self.sa = SetAbstraction(npoint=K, K=16, in_channel=0, mlp=[32, 64, 128], bn=False)
        self.pn = PointNet(in_channel=3 + 128, mlps=[128, 256, 512, d], relu=[True, True, True, False], bn=False)
        self.inv_pool = nn.Sequential(nn.Linear(d, 256), nn.ReLU(), nn.Linear(256, 1024), nn.ReLU(),
                                      nn.Linear(1024, k * 128), nn.ReLU())
        self.inv_mlp = MLP(in_channel=d + 128, mlps=[128, 64, 32, 3], relu=[True, True, True, False], bn=False)
        self.K, self.k, self.d, self.L = K, k, d, L
        self.quantize = STEQuantize.apply

    # -- the two halves the scripts use separately (compress.py:113-127, decompress.py:96-102) --
    def encode_patches(self, patches):
        """patches [BS, K, 3] (recentred, scaled) -> (latent [BS, d] after the sigmoid spread, rounded latent)."""
        feat = self.sa.forward_points(patches, out_dtype=torch.bfloat16)          # AE.py:38
        latent = self.pn.forward_xyz_feat(patches, feat)                          # AE.py:39
        # AE.py:42-45 in one kernel; it also emits the rounded latent as zero-padded bf16 rows, the operand of inv_pool's first
        # GEMM, which decode_patches picks up when it is handed this very tensor (compress -> decompress in one process)
        latent, latent_q, qb = ops.quantise_latent(latent, self.L - 0.2, kpad=(self.d + 63) // 64 * 64)
        self._q_pad = (latent_q, latent_q._version, qb)   # holds the tensor itself: its storage cannot be recycled under the key
        return latent, latent_q

    def decode_patches(self, latent_q):
        """latent_q [BS, d] -> patches [BS, k, 3]   (AE.py:48-53)."""
        BS = latent_q.shape[0]
        # inv_pool's last Linear emits [128 channels, k points] per patch (AE.py:49 `view(BS, -1, k)`); permuting its
        # rows once makes the GEMM write [k points, 128 channels] (channel-last) directly.
        w4, b4 = self.inv_pool[4].weight, self.inv_pool[4].bias
        key = (w4.data_ptr(), w4._version, b4._version)
        if getattr(self, "_perm_key", None) != key:
            self._perm_w4 = w4.detach().view(128, self.k, -1).permute(1, 0, 2).reshape(128 * self.k, -1).contiguous()
            self._perm_b4 = b4.detach().view(128, self.k).t().reshape(-1).contiguous()
            self._perm_key = key
        inv_layers = [(self.inv_pool[0].weight, self.inv_pool[0].bias, True), (self.inv_pool[2].weight, self.inv_pool[2].bias, True),
                      (self._perm_w4, self._perm_b4, True)]
        if all(mlp_ops.linear_supported(BS, w.shape[0]) for w, _, _ in inv_layers):
            # AE.py:19-26 on the streamed tensor-core GEMM (csrc/gemm_ws.cu); the latent is zero padded to the K granule
            cached = getattr(self, "_q_pad", None)
            base = latent_q._base if latent_q._base is not None else latent_q
            if (cached is not None and base is cached[0] and latent_q._version == cached[1] and latent_q.is_contiguous() and
                    latent_q.numel() == cached[0].numel() and latent_q.data_ptr() == cached[0].data_ptr()):
                lat = cached[2]
            else:
                lat = torch.nn.functional.pad(latent_q.detach().to(torch.bfloat16), (0, (-self.d) % 64))
            lin = mlp_ops.stream_chain(lat, inv_layers)
        else:
            lin = mlp_ops.library_chain(latent_q.detach(), inv_layers, out_dtype=torch.bfloat16)
        # AE.py:50-52: cat(features, tiled latent) -> inv_mlp, as two input segments of the fused chain
        out = mlp_ops.fused_chain([(lin.view(BS * self.k, 128), 1), (latent_q.detach().float().contiguous(), self.k)],
                                  self.inv_mlp.layers())
        return out.view(BS, self.k, 3)

    def forward_train(self, xyz):
        """AE.forward (AE.py:34-55) under autograd: fp32 library GEMMs for the network bodies, pcc kernels for the
        grouping; gradients reach every parameter through the STE quantiser exactly as in the reference."""
        BS = xyz.shape[0]
        feat = self.sa.forward_points_train(xyz)
        h = torch.cat((xyz, feat), dim=2).reshape(BS * xyz.shape[1], -1)          # AE.py:39
        for w, b, relu in self.pn.layers():
            h = torch.addmm(b, h, w.t())
            if relu:
                h = torch.relu(h)
        latent = h.view(BS, xyz.shape[1], -1).max(dim=1)[0]
        spread = self.L - 0.2
        latent = torch.sigmoid(latent) * spread - spread / 2                      # AE.py:42-44
        latent_q = STEQuantize.apply(latent)                                      # AE.py:45
        lin = self.inv_pool(latent_q).view(BS, -1, self.k)                        # AE.py:48-49  [BS,128,k]
        x = torch.cat((lin, latent_q.unsqueeze(-1).repeat((1, 1, self.k))), dim=1)  # AE.py:50-51 [BS,144,k]
        h = x.permute(0, 2, 1).reshape(BS * self.k, -1)
        for w, b, relu in self.inv_mlp.layers():
            h = torch.addmm(b, h, w.t())
            if relu:
                h = torch.relu(h)
        return h.view(BS, self.k, 3), latent, latent_q

    def forward(self, xyz):
        """Reference signature (AE.py:34-55): xyz [BS, K, 3] -> (new_xyz [BS, k, 3], latent, latent_quantized).
        With autograd enabled and trainable parameters the differentiable body runs; otherwise the fused inference path."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self.forward_train(xyz.contiguous())
        latent, latent_q = self.encode_patches(xyz.contiguous())
        return self.decode_patches(latent_q), latent, latent_q


class ConditionalProbabilityModel(nn.Module):
    """AE.ConditionalProbabilityModel(L, d) (/root/reference/AE.py:87-123): PMF of every latent symbol given the patch
    centres.  Tiny (0.5 M parameters, 63 MFLOP per cloud): plain library GEMMs, same parameter keys."""

    def __init__(self, L, d):
        super().__init__()
        self.L, self.d = L, d
        self.model_pn = PointNet(in_channel=3, mlps=[64, 128, 256], relu=[True, True, True], bn=False)
        self.model_mlp = nn.Sequential(nn.Conv2d(3 + 256, 512, 1), nn.ReLU(), nn.Conv2d(512, 512, 1), nn.ReLU(),
                                       nn.Conv2d(512, d * L, 1))

    def forward(self, sampled_xyz):
        B, S, _ = sampled_xyz.shape
        h = sampled_xyz.reshape(B * S, 3)
        for w, b, relu in self.model_pn.layers():
            h = torch.relu(torch.addmm(b, h, w.t())) if relu else torch.addmm(b, h, w.t())
        feature = h.view(B, S, -1).max(dim=1)[0]                                  # AE.py:112
        x = torch.cat((sampled_xyz, feature[:, None, :].expand(-1, S, -1)), dim=2).reshape(B * S, -1)  # AE.py:115
        for i in (0, 2, 4):
            m = self.model_mlp[i]
            x = torch.addmm(m.bias, x, m.weight.flatten(1).t())
            if i < 4:
                x = torch.relu(x)
        return torch.softmax(x.view(B, S, self.d, self.L), dim=3)                 # AE.py:118-121
