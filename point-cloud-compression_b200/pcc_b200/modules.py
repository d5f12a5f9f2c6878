"""Parameter containers with the reference's state_dict keys, and the batched device forward of the IPDAE patch
auto-encoder (/root/reference/AE.py:12-55) built on the pcc_b200 kernels.

The containers mirror the reference classes only as far as names, constructor arguments and parameter keys /
shapes go, so `ae.pkl` checkpoints written by the reference's train.py load unchanged:
    sa.conv{0,1,2}.{weight,bias}                 pn_kit.SetAbstraction   (pn_kit.py:146-162)
    pn.mlp_Modules.{0..3}.0.{weight,bias}        pn_kit.PointNet         (pn_kit.py:98-121)
    inv_pool.{0,2,4}.{weight,bias}               AE.inv_pool             (AE.py:19-26)
    inv_mlp.mlp_Modules.{0..3}.0.{weight,bias}   pn_kit.MLP              (pn_kit.py:263-287)
The forward bodies are NOT the reference's: they run channel-last on flattened [rows, C] activations and call the
fused kernels (in-patch kNN -> shared MLP -> max over neighbours, ...).  The bodies themselves live in bodies.py, written
against the reference's attribute names, so the same code serves these containers and -- through install() -- instances of
the reference's own classes.
"""
import torch
import torch.nn as nn

from . import bodies, ops


class STEQuantize(torch.autograd.Function):
    """Round in the forward pass, identity gradient (AE.STEQuantize, /root/reference/AE.py:72-85)."""

    @staticmethod
    def forward(ctx, x):
        return x.round()

    @staticmethod
    def backward(ctx, grad_outputs):
        return grad_outputs


def _conv_stack(channels, relu):
    """pn_kit.py:104-121: one nn.Sequential(Conv2d[, ReLU]) per layer -- keys "<i>.0.weight" / "<i>.0.bias"."""
    mods = nn.ModuleList()
    for cin, cout, r in zip(channels[:-1], channels[1:], relu):
        mods.append(nn.Sequential(nn.Conv2d(cin, cout, 1), nn.ReLU()) if r else nn.Sequential(nn.Conv2d(cin, cout, 1)))
    return mods


class SetAbstraction(nn.Module):
    """pn_kit.SetAbstraction(npoint, K, in_channel, mlp, bn=False, finalRelu=True) with S == N (no FPS)."""

    def __init__(self, npoint, K, in_channel, mlp, bn=False, finalRelu=True):
        super().__init__()
        if bn:
            raise NotImplementedError("pcc_b200.SetAbstraction: bn=True is not used by the reference AE")
        self.npoint, self.K, self.finalRelu, self.bn = npoint, K, finalRelu, False
        self.conv0 = nn.Conv2d(in_channel + 3, mlp[0], 1)
        self.conv1 = nn.Conv2d(mlp[0], mlp[1], 1)
        self.conv2 = nn.Conv2d(mlp[1], mlp[2], 1)

    def layers(self):
        return [(self.conv0.weight.flatten(1), self.conv0.bias, True), (self.conv1.weight.flatten(1), self.conv1.bias, True),
                (self.conv2.weight.flatten(1), self.conv2.bias, self.finalRelu)]

    def forward_points(self, xyz, out_dtype=torch.float32):
        """xyz [BS, P, 3] (channel-last) -> per-point features [BS, S, C_out] (channel-last).
        pn_kit.py:164-211: kNN(K) in the patch, recentre on the query, shared MLP, max over the K neighbours."""
        return bodies.sa_points(self, xyz, out_dtype)[1]

    def forward_points_train(self, xyz):
        """fp32 torch body under autograd (library GEMMs) -- part of AE.forward_train_fp32, the arithmetic the kernel training
        path is tested against: the kNN grouping runs on the pcc kernel (no gradient flows into it)."""
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
        new_xyz, feat = bodies.sa_points(self, xyz.permute(0, 2, 1).contiguous())
        return new_xyz.permute(0, 2, 1), feat.permute(0, 2, 1)


class PointNet(nn.Module):
    """pn_kit.PointNet(in_channel, mlps, relu, bn): shared MLP then max over the points."""

    def __init__(self, in_channel, mlps, relu, bn=False):
        super().__init__()
        if bn:
            raise NotImplementedError("pcc_b200.PointNet: bn=True is not used by the reference AE")
        self.relu = list(relu)
        self.mlp_Modules = _conv_stack([in_channel] + list(mlps), self.relu)

    def layers(self):
        return [(m[0].weight.flatten(1), m[0].bias, r) for m, r in zip(self.mlp_Modules, self.relu)]

    def forward_points(self, x):
        """x [BS, P, C] channel-last -> [BS, D]."""
        return bodies.pointnet_points(self, x)

    def forward_xyz_feat(self, xyz, feat):
        """The AE.py:39 call `pn(cat((xyz, feat)))` without materialising the concatenation (bodies.pointnet_xyz_feat)."""
        return bodies.pointnet_xyz_feat(self, xyz, feat)

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
        self.mlp_Modules = _conv_stack([in_channel] + list(mlps), self.relu)

    def layers(self):
        return [(m[0].weight.flatten(1), m[0].bias, r) for m, r in zip(self.mlp_Modules, self.relu)]

    def forward_points(self, x):
        return bodies.mlp_points(self, x)

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
        return bodies.ae_encode(self, patches)

    def decode_patches(self, latent_q):
        """latent_q [BS, d] -> patches [BS, k, 3]   (AE.py:48-53)."""
        return bodies.ae_decode(self, latent_q)

    def forward_train(self, xyz):
        """AE.forward (AE.py:34-55) under autograd, forward and backward contractions on the pcc kernels (bodies.ae_forward_train:
        bf16 operands, fp32 accumulation and weight gradients)."""
        return bodies.ae_forward_train(self, xyz)

    def forward_train_fp32(self, xyz):
        """The same body as plain fp32 torch ops under autograd (library GEMMs): the reference arithmetic the kernel path is
        tested against (tests/test_gpu_train.py); Trainer(kernels=False) trains with it."""
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
        if bodies.training_pass(self):
            return self.forward_train(xyz.contiguous())
        return bodies.ae_forward(self, xyz)


class ConditionalProbabilityModel(nn.Module):
    """AE.ConditionalProbabilityModel(L, d) (/root/reference/AE.py:87-123): PMF of every latent symbol given the patch
    centres, same parameter keys.  Inference runs on the pcc kernels (bodies.prob_forward: batch-invariant, so a stream coded
    in a batch of 32 decodes in a batch of 1); under autograd the differentiable body below runs."""

    def __init__(self, L, d):
        super().__init__()
        self.L, self.d = L, d
        self.model_pn = PointNet(in_channel=3, mlps=[64, 128, 256], relu=[True, True, True], bn=False)
        self.model_mlp = nn.Sequential(nn.Conv2d(3 + 256, 512, 1), nn.ReLU(), nn.Conv2d(512, 512, 1), nn.ReLU(),
                                       nn.Conv2d(512, d * L, 1))

    def forward_train(self, sampled_xyz):
        return bodies.prob_forward_train(self, sampled_xyz)

    def forward_train_fp32(self, sampled_xyz):
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

    def forward(self, sampled_xyz):
        if bodies.training_pass(self):
            return self.forward_train(sampled_xyz)
        return bodies.prob_forward(self, sampled_xyz)
