"""Parameter containers with the reference's state_dict keys for the PointNet++ encoder of pppe_pcd_ae.py
(/root/reference/pppe_pcd_ae.py:573-690: PointNetSetAbstraction, PointNetSetAbstractionMSG, PointNet2EncoderFull) and its fused
device forward -- the "fast pppe_pcd_ae compress" leg of BASELINE cfg5.  Keys:
    sa_modules.0.branches.{j}.mlp_stack.{l}.0.weight / .1.{weight,bias,running_mean,running_var,num_batches_tracked}   (MSG level)
    sa_modules.{i}.mlp_stack.{l}.0.weight / .1.*                                                                        (SS levels)
    global_conv.0.weight, global_conv.1.*, global_conv.3.{weight,bias}
The bodies (bodies.pppe_*) read the parameters by the reference's attribute names, so install() binds the same code to the
reference's own classes.  Inference only (eval-mode BatchNorm folded into the convolutions, no autograd graph); a train-mode call raises.
"""
import torch
import torch.nn as nn

from . import bodies


def _conv2d_bn_relu(cin, cout):
    return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


def _inference_only(mod):
    if bodies.has_train_mode_bn(mod):
        raise NotImplementedError("pcc_b200.pppe: inference only (eval-mode BatchNorm is folded into the convolutions) -- call "
                                  ".eval(); training the pppe encoder stays on the reference's own modules")


class PointNetSetAbstraction(nn.Module):
    """pppe_pcd_ae.PointNetSetAbstraction(npoint, K, in_channel, mlp, bn=True)."""

    def __init__(self, npoint, K, in_channel, mlp, bn=True):
        super().__init__()
        self.npoint, self.K = npoint, K
        last, layers = in_channel + 3, []
        for out in mlp:
            layers.append(_conv2d_bn_relu(last, out) if bn else nn.Sequential(nn.Conv2d(last, out, 1), nn.ReLU()))
            last = out
        self.mlp_stack = nn.ModuleList(layers)

    def forward(self, xyz, points=None):
        _inference_only(self)
        with torch.no_grad():
            return bodies.pppe_sa_forward(self, xyz, points)


class PointNetSetAbstractionMSG(nn.Module):
    """pppe_pcd_ae.PointNetSetAbstractionMSG(npoint, scales, in_channel, bn=True)."""

    def __init__(self, npoint, scales, in_channel, bn=True):
        super().__init__()
        self.branches = nn.ModuleList(PointNetSetAbstraction(npoint, sc["K"], in_channel, sc["mlp"], bn=bn) for sc in scales)

    def forward(self, xyz, points=None):
        _inference_only(self)
        with torch.no_grad():
            return bodies.pppe_sa_forward(self, xyz, points)


DEFAULT_SA_BLOCKS = (   # pppe_pcd_ae.py:642-646
    {"type": "MSG", "npoint": 512, "scales": [{"K": 16, "mlp": [32, 32, 64]}, {"K": 32, "mlp": [64, 64, 128]}], "in_channel": 0},
    {"type": "SS", "npoint": 128, "K": 32, "mlp": [128, 128, 256], "in_channel": 64 + 128},
    {"type": "SS", "npoint": 32, "K": 32, "mlp": [256, 256, 512], "in_channel": 256},
)


class PointNet2EncoderFull(nn.Module):
    """pppe_pcd_ae.PointNet2EncoderFull(sa_blocks=None, latent_dim=256, bn=True): x [B, N, 3] -> (latent, pooled features)."""

    def __init__(self, sa_blocks=None, latent_dim=256, bn=True):
        super().__init__()
        blocks = list(DEFAULT_SA_BLOCKS if sa_blocks is None else sa_blocks)
        mods = []
        for blk in blocks:
            if blk["type"] == "MSG":
                mods.append(PointNetSetAbstractionMSG(blk["npoint"], blk["scales"], blk.get("in_channel", 0), bn=bn))
            else:
                mods.append(PointNetSetAbstraction(blk["npoint"], blk["K"], blk.get("in_channel", 0), blk["mlp"], bn=bn))
        self.sa_modules = nn.ModuleList(mods)
        last = blocks[-1]
        out_c = sum(s["mlp"][-1] for s in last["scales"]) if last["type"] == "MSG" else last["mlp"][-1]
        self.global_conv = nn.Sequential(nn.Conv1d(out_c, out_c, 1, bias=False), nn.BatchNorm1d(out_c), nn.ReLU(inplace=True),
                                         nn.Conv1d(out_c, latent_dim, 1))
        self.latent_dim = latent_dim

    def forward(self, x):
        _inference_only(self)
        with torch.no_grad():
            return bodies.pppe_encoder_forward(self, x)


def quantize_st(x, min_val, max_val, levels):
    """pppe_pcd_ae.quantize_st (pppe_pcd_ae.py:721-737), forward values."""
    scaled = (torch.clamp(x, min_val, max_val) - min_val) / (max_val - min_val + 1e-9) * (levels - 1)
    return torch.clamp(torch.round(scaled), 0, levels - 1)


def compress(encoder, x, latent_bins=7):
    """The encoder half of pppe_pcd_ae.PointCloudAE.forward (pppe_pcd_ae.py:866-872): x [B, N, 3] -> (quantised latent [B, d]
    in [0, latent_bins - 1], pooled conditioning features [B, C_out]).  (The reference tiles the latent to [B, d, N] before
    quantising; every column is the same value, so one column is kept.)"""
    latent, cond = encoder(x)
    return quantize_st(latent, 0.0, latent_bins - 1.0, latent_bins), cond
