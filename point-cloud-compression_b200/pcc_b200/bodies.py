"""Fused inference bodies of the reference's network modules, written against the reference's ATTRIBUTE NAMES only.

Every function takes a module instance `mod` -- either one of pcc_b200's own parameter containers (modules.py, pppf.py) or an
instance of the reference's own class (pn_kit.SetAbstraction / PointNet / MLP, AE.AE, AE.ConditionalProbabilityModel,
pointnet_sa_module.PointnetSAModule, PPPF_AE.FoldingNet / PPPF_AE / ConditionalProbabilityModel) -- reads its parameters
through the names the reference gives them and runs the batched device forward on the pcc kernels.
`install.patch_reference_modules()` binds these as the `forward` of the reference classes, so the unmodified scripts
(compress.py, decompress.py, eval.py; train.py in its no-grad sections) reach the tcgen05 kernels.

There is no library GEMM and no CPU path in here: a shape the kernels do not take raises (PCC_ERR_UNSUPPORTED ->
ValueError).  Element-wise glue (sigmoid, round, softmax, max over a tiny axis) is torch.
"""
import os

import torch
import torch.nn as nn

from . import mlp_ops, ops, pn_kit_ops

_BN_TYPES = (nn.BatchNorm1d, nn.BatchNorm2d)
_TWO_LAUNCH_PN = bool(os.environ.get("PCC_PN_TWO_LAUNCH"))   # A/B switch: PointNet front + tail as two launches (round 1)
_GROUPED_SA = bool(os.environ.get("PCC_SA_GROUPED"))   # A/B switch for the measurements in profiles/: the general (grouped tensor) route


def training_pass(mod):
    """True when the call must stay differentiable: autograd on and trainable parameters (the reference's train loop)."""
    return torch.is_grad_enabled() and any(p.requires_grad for p in mod.parameters())


def has_train_mode_bn(mod):
    return mod.training and any(isinstance(m, _BN_TYPES) for m in mod.modules())


def _state_key(mod):
    # (storage, version) of every parameter / buffer + the epoch that mlp_ops.invalidate_weight_caches() bumps (updates that do
    # not touch the version counters, e.g. torch's fused optimiser kernels)
    return (mlp_ops._weights_epoch,) + tuple((t.data_ptr(), t._version) for t in list(mod.parameters()) + list(mod.buffers()))


def _cached(mod, name, build):
    """Per-instance cache of derived tensors (folded / rotated / permuted weights), rebuilt when any parameter or buffer of
    the module changes.  The cached tensors keep their identity between calls, which is what mlp_ops' pack caches (and CUDA
    graph captures) key on."""
    store = mod.__dict__.setdefault("_pcc_cache", {})
    key = _state_key(mod)
    hit = store.get(name)
    if hit is None or hit[0] != key:
        with torch.no_grad():
            hit = store[name] = (key, build())
    return hit[1]


def _w2d(conv):
    return conv.weight.flatten(1)


def _fold(conv, bn):
    """Eval-mode BatchNorm folded into the preceding 1x1 convolution: y = (Wx + b - mean) * gamma / sqrt(var + eps) + beta."""
    w = _w2d(conv)
    b = conv.bias if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    return (w * scale[:, None]).contiguous(), ((b - bn.running_mean) * scale + bn.bias).contiguous()


def _seq_layer(seq):
    """One nn.Sequential(conv[, bn][, relu]) block of pn_kit.PointNet / MLP (pn_kit.py:104-121) -> (w, b, relu)."""
    mods = list(seq)
    conv = mods[0]
    bn = next((m for m in mods[1:] if isinstance(m, _BN_TYPES)), None)
    relu = any(isinstance(m, nn.ReLU) for m in mods[1:])
    if bn is not None:
        return _fold(conv, bn) + (relu,)
    return _w2d(conv), conv.bias, relu


def stack_layers(mod):
    """pn_kit.PointNet / pn_kit.MLP: `mlp_Modules` (pn_kit.py:98-121, 263-287)."""
    return _cached(mod, "stack", lambda: [_seq_layer(s) for s in mod.mlp_Modules])


def sa_layers(mod):
    """pn_kit.SetAbstraction: conv0..2 (+ bn0..2), ReLU after the first two and, with finalRelu, the third (pn_kit.py:198-205)."""
    def build():
        out = []
        for i in range(3):
            conv = getattr(mod, f"conv{i}")
            relu = True if i < 2 else bool(mod.finalRelu)
            if getattr(mod, "bn", False):
                out.append(_fold(conv, getattr(mod, f"bn{i}")) + (relu,))
            else:
                out.append((_w2d(conv), conv.bias, relu))
        return out
    return _cached(mod, "sa", build)


def triple_layers(mod):
    """pointnet_sa_module.PointnetSAModule.mlp: Conv2d + BatchNorm2d + ReLU triples (pointnet_sa_module.py:49-56)."""
    def build():
        mods = list(mod.mlp)
        return [_fold(mods[i], mods[i + 1]) + (True,) for i in range(0, len(mods), 3)]
    return _cached(mod, "triples", build)


# ---- pn_kit.SetAbstraction / PointNet / MLP (channel-last bodies) ----------------------------------------------------------
def sa_points(mod, xyz, out_dtype=torch.float32):
    """pn_kit.SetAbstraction.forward (pn_kit.py:164-211) on channel-last xyz [BS, P, 3]: (new_xyz [BS, S, 3], per-point
    features [BS, S, C_out]) -- kNN(K) around every query, recentre, shared MLP, max over the K neighbours."""
    BS, P, _ = xyz.shape
    S = mod.npoint
    if S == P:                                                                     # pn_kit.py:181-182
        new_xyz = xyz
    else:                                                                          # pn_kit.py:184 (CPU-RNG start index)
        new_xyz = ops.gather(xyz, pn_kit_ops.farthest_point_sample_batch(xyz, S))
    layers = sa_layers(mod)
    if S == P and BS <= 65535 and mlp_ops.sa_indexed_supported(P, mod.K, layers) and not _GROUPED_SA:
        # the AE's shape: kNN table as bytes, then the chain gathers the recentred neighbours from the patch itself -- the
        # [BS, S, K, 3] grouped tensor is never written (same fp32 arithmetic: bit-identical to the general route below)
        feat = mlp_ops.sa_chain_indexed(xyz, ops.knn_patch_u8(xyz, mod.K), layers, out_dtype=out_dtype)     # :190-207
        return new_xyz, feat.reshape(BS, S, -1)
    _, _, grouped = ops.knn(new_xyz, xyz, mod.K, return_nn=True, centre_sub=True, nn_only=True)  # :190-191  [BS,S,K,3]
    feat = mlp_ops.fused_chain(grouped.reshape(BS * S * mod.K, 3), layers, group=mod.K, out_dtype=out_dtype)
    return new_xyz, feat.reshape(BS, S, -1)


def pointnet_points(mod, x):
    """pn_kit.PointNet.forward (pn_kit.py:124-144) on channel-last x [BS, P, C] -> [BS, D]."""
    BS, P, C = x.shape
    return mlp_ops.run_chain(x.reshape(BS * P, C), stack_layers(mod), group=P)


def pointnet_xyz_feat(mod, xyz, feat):
    """The AE.py:39 call `pn(cat((xyz, feat)))` without materialising the concatenation: xyz [BS,P,3] fp32 and feat [BS,P,F]
    (bf16 from the SetAbstraction kernel) are two input segments of the fused chain; the first layer's weight columns are
    rotated once so the 16-byte aligned feature block comes first."""
    BS, P, F = feat.shape
    layers = list(stack_layers(mod))

    def rot():
        w0 = layers[0][0]
        return torch.cat((w0[:, 3:], w0[:, :3]), dim=1).detach().contiguous()

    layers[0] = (_cached(mod, "rot_w0", rot), layers[0][1], layers[0][2])
    f2, x2 = feat.reshape(BS * P, F), xyz.reshape(BS * P, 3)
    if not _TWO_LAUNCH_PN and mlp_ops.pointnet_fused_supported(f2, x2, layers, P):
        return mlp_ops.pointnet_fused(f2, x2, layers)                              # one launch, no [M, 256] activation in HBM
    return mlp_ops.run_chain([(f2, 1), (x2, 1)], layers, group=P)


def mlp_points(mod, x):
    """pn_kit.MLP.forward (pn_kit.py:289-305) on channel-last x [BS, P, C] -> [BS, P, D]."""
    BS, P, C = x.shape
    return mlp_ops.run_chain(x.reshape(BS * P, C), stack_layers(mod)).reshape(BS, P, -1)


# ---- AE.AE -----------------------------------------------------------------------------------------------------------------
def ae_latent_dim(mod):
    return stack_layers(mod.pn)[-1][0].shape[0]


def ae_encode(mod, patches):
    """AE.py:37-45: patches [BS, K, 3] (recentred, scaled) -> (latent [BS, d] after the sigmoid spread, rounded latent)."""
    _, feat = sa_points(mod.sa, patches, out_dtype=torch.bfloat16)                   # AE.py:38
    raw = pointnet_xyz_feat(mod.pn, patches, feat)                                  # AE.py:39
    d = raw.shape[1]
    # AE.py:42-45 in one kernel; it also emits the rounded latent as zero-padded bf16 rows, the operand of inv_pool's first
    # GEMM, which ae_decode picks up when it is handed this very tensor (compress -> decompress in one process)
    latent, latent_q, qb = ops.quantise_latent(raw, mod.L - 0.2, kpad=(d + 63) // 64 * 64)
    mod.__dict__["_q_pad"] = (latent_q, latent_q._version, qb)   # holds the tensor itself: its storage cannot be recycled
    return latent, latent_q


def ae_decode(mod, latent_q):
    """AE.py:48-53: latent_q [BS, d] -> patches [BS, k, 3]."""
    BS, d = latent_q.shape
    k = mod.k
    l0, l2, l4 = mod.inv_pool[0], mod.inv_pool[2], mod.inv_pool[4]

    # inv_pool's last Linear emits [128 channels, k points] per patch (AE.py:49 `view(BS, -1, k)`); permuting its rows once
    # makes the GEMM write [k points, 128 channels] (channel-last) directly.
    def perm():
        ch = l4.weight.shape[0] // k
        return (l4.weight.detach().view(ch, k, -1).permute(1, 0, 2).reshape(ch * k, -1).contiguous(),
                l4.bias.detach().view(ch, k).t().reshape(-1).contiguous())

    w4, b4 = _cached(mod.inv_pool, "perm_w4", perm)
    inv_layers = [(l0.weight, l0.bias, True), (l2.weight, l2.bias, True), (w4, b4, True)]
    cached = mod.__dict__.get("_q_pad")
    base = latent_q._base if latent_q._base is not None else latent_q
    if (cached is not None and base is cached[0] and latent_q._version == cached[1] and latent_q.is_contiguous() and
            latent_q.numel() == cached[0].numel() and latent_q.data_ptr() == cached[0].data_ptr()):
        lat = cached[2]
    else:
        lat = torch.nn.functional.pad(latent_q.detach().to(torch.bfloat16), (0, (-d) % 64))
    lin = mlp_ops.stream_chain(lat, inv_layers)                                    # AE.py:19-26,48 on csrc/gemm_ws.cu
    ch = w4.shape[0] // k
    # AE.py:50-52: cat(features, tiled latent) -> inv_mlp, as two input segments of the fused chain
    out = mlp_ops.run_chain([(lin.view(BS * k, ch), 1), (latent_q.detach().float().contiguous(), k)], stack_layers(mod.inv_mlp))
    return out.view(BS, k, -1)


def linear_stack_forward(seq, x):
    """An nn.Sequential of Linear (+ ReLU) layers applied to [rows, d] -- AE.inv_pool called on its own, as decompress.py:96 does
    (`ae.inv_pool(latent_quantized)`): every Linear on the streamed tensor-core GEMM, reference layout [rows, out_features],
    fp32 dtype (the values are the bf16 activations the fused decode uses)."""
    def build():
        mods, out = list(seq), []
        for i, m in enumerate(mods):
            if isinstance(m, nn.Linear):
                bias = m.bias if m.bias is not None else torch.zeros(m.out_features, device=m.weight.device)
                out.append((m.weight, bias, i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)))
            elif not isinstance(m, nn.ReLU):
                raise ValueError(f"pcc_b200: {type(m).__name__} inside a Linear stack is not a layer the kernels take")
        return out
    layers = _cached(seq, "linear_stack", build)
    d = x.shape[-1]
    lat = torch.nn.functional.pad(x.detach().reshape(-1, d).to(torch.bfloat16), (0, (-d) % 64))
    out = mlp_ops.stream_chain(lat, layers)
    return out.float().reshape(x.shape[:-1] + (out.shape[-1],))


def ae_forward(mod, xyz):
    """AE.AE.forward (AE.py:34-55): xyz [BS, K, 3] -> (new_xyz [BS, k, 3], latent, latent_quantized)."""
    latent, latent_q = ae_encode(mod, xyz.contiguous())
    return ae_decode(mod, latent_q), latent, latent_q


# ---- training bodies: the same modules under autograd, every contraction (forward and backward) on the pcc kernels --------------
def _grad_layers_stack(mod):
    """(w, b, relu) of pn_kit.PointNet / MLP with autograd-tracked views of the parameters (no BatchNorm: the AE uses bn=False)."""
    out = []
    for seq in mod.mlp_Modules:
        mods = list(seq)
        if any(isinstance(m, _BN_TYPES) for m in mods):
            raise NotImplementedError("pcc_b200 training bodies: BatchNorm inside a trained stack is not supported")
        out.append((mods[0].weight.flatten(1), mods[0].bias, any(isinstance(m, nn.ReLU) for m in mods[1:])))
    return out


def ae_forward_train(mod, xyz):
    """AE.AE.forward (AE.py:34-55) under autograd: xyz [BS, K, 3] -> (new_xyz [BS, k, 3], latent, latent_quantized); gradients reach
    every parameter through the STE quantiser exactly as in the reference.  The grouping (in-patch kNN) carries no gradient."""
    from . import train_ops as T
    if getattr(mod.sa, "bn", False):
        raise NotImplementedError("pcc_b200 training bodies: SetAbstraction(bn=True) is not supported")
    xyz = xyz.contiguous()
    BS, P, _ = xyz.shape
    K = mod.sa.K
    sa_l = [(getattr(mod.sa, f"conv{i}").weight.flatten(1), getattr(mod.sa, f"conv{i}").bias, True if i < 2 else bool(mod.sa.finalRelu))
            for i in range(3)]
    fused = (K == 16 and P <= 256 and P % 8 == 0 and sa_l[2][2] and [tuple(w.shape) for w, _, _ in sa_l] == [(32, 3), (64, 32), (128, 64)]
             and all(b is not None for _, b, _ in sa_l) and not os.environ.get("PCC_SA_TRAIN_UNFUSED"))
    if fused:
        # one forward kernel (the inference one: no activation is kept) and one backward kernel that recomputes them per tile
        with torch.no_grad():
            idx8 = ops.knn_patch_u8(xyz, K)                                                            # pn_kit.py:190
        feat = T.sa_indexed_train(xyz, idx8, sa_l, out_bf16=True)                                      # pn_kit.py:191-207  bf16 [BS*P, 128]
    else:
        with torch.no_grad():
            _, _, grouped = ops.knn(xyz, xyz, K, return_nn=True, centre_sub=True, nn_only=True)        # pn_kit.py:190-191
        x1 = T.fold_first_train(grouped.reshape(BS * P * K, 3), sa_l[0][0], sa_l[0][1])               # conv0 in fp32 on the CUDA cores
        feat = T.mlp_train(x1, sa_l[1:], group=K, mode="pool", x0_is_relu=True)                       # pn_kit.py:196-207  [BS*P, 128]
    pn_l = _grad_layers_stack(mod.pn)
    w0 = pn_l[0][0]
    pn_l[0] = (torch.cat((w0[:, 3:], w0[:, :3]), dim=1), pn_l[0][1], pn_l[0][2])                      # AE.py:39 cat(xyz, feat) -> [feat | xyz]
    F_ = feat.shape[1]
    x0 = torch.cat((feat.to(torch.bfloat16), xyz.reshape(BS * P, 3).to(torch.bfloat16),
                    torch.zeros((BS * P, (-(F_ + 3)) % 64), dtype=torch.bfloat16, device=xyz.device)), dim=1)
    raw = T.mlp_train(x0, pn_l, group=P, mode="pool")                                                  # pn_kit.py:124-144  [BS, d]
    spread = mod.L - 0.2
    latent = torch.sigmoid(raw) * spread - spread / 2                                                  # AE.py:42-44
    latent_q = mod.quantize(latent)                                                                    # AE.py:45 (STE)
    k = mod.k
    ip = [(mod.inv_pool[i].weight, mod.inv_pool[i].bias, True) for i in (0, 2, 4)]
    # AE.py:49 views the last Linear's output as [channels, k] per patch and permutes it to [k, channels]; permuting that layer's
    # ROWS instead (as ae_decode does for inference) moves 2 x 67 MB of weights per step where the activation and its gradient
    # were 2 x 2 x 67 MB of strided copies, and the GEMM writes the layout the decoder reads
    l4 = mod.inv_pool[4]
    ch = l4.out_features // k
    if os.environ.get("PCC_TRAIN_PERMUTE_ACT"):   # A/B: the reference's own order of operations
        lin = T.mlp_train(T.pad_bf16(latent_q, 64), ip, mode="bf16")                                    # AE.py:48  bf16 [BS, 128 * k]
        lin_pts = lin.view(BS, ch, k).permute(0, 2, 1).reshape(BS * k, ch)                              # AE.py:49
    else:
        ip[2] = (l4.weight.view(ch, k, -1).permute(1, 0, 2).reshape(ch * k, -1), l4.bias.view(ch, k).t().reshape(-1), True)
        lin = T.mlp_train(T.pad_bf16(latent_q, 64), ip, mode="bf16")                                    # AE.py:48  bf16 [BS, k * 128]
        lin_pts = lin.view(BS * k, ch)                                                                  # AE.py:49
    d = latent_q.shape[1]
    x0d = torch.cat((lin_pts, latent_q.to(torch.bfloat16).repeat_interleave(k, dim=0),
                     torch.zeros((BS * k, (-(ch + d)) % 64), dtype=torch.bfloat16, device=xyz.device)), dim=1)   # AE.py:50-51
    out = T.mlp_train(x0d, _grad_layers_stack(mod.inv_mlp), mode="f32")                                 # AE.py:52
    return out.view(BS, k, 3), latent, latent_q


def prob_forward_train(mod, sampled_xyz):
    """AE.ConditionalProbabilityModel.forward (AE.py:107-123) under autograd on the pcc kernels."""
    from . import train_ops as T
    B, S, _ = sampled_xyz.shape
    xyz = sampled_xyz.detach().float().contiguous()
    feature = T.mlp_train(T.pad_bf16(xyz.reshape(B * S, 3), 64), _grad_layers_stack(mod.model_pn), group=S, mode="pool")   # AE.py:112
    m0, m2, m4 = mod.model_mlp[0], mod.model_mlp[2], mod.model_mlp[4]
    F_ = feature.shape[1]
    x0 = torch.cat((xyz.reshape(B * S, 3).to(torch.bfloat16), feature.to(torch.bfloat16).repeat_interleave(S, dim=0),
                    torch.zeros((B * S, (-(F_ + 3)) % 64), dtype=torch.bfloat16, device=xyz.device)), dim=1)                # AE.py:115
    logits = T.mlp_train(x0, [(m0.weight.flatten(1), m0.bias, True), (m2.weight.flatten(1), m2.bias, True),
                              (m4.weight.flatten(1), m4.bias, False)], mode="f32")                                        # AE.py:116-118
    return torch.softmax(logits.view(B, S, mod.d, mod.L), dim=3)                                                           # AE.py:119-121


# ---- conditional probability models ---------------------------------------------------------------------------------------
def _prob_tail(model_mlp, sampled_xyz, feature, d, L):
    """AE.py:115-121 / PPPF_AE.py:213-228: cat(xyz, tiled global feature) -> Conv2d 3+F -> 512 -> 512 -> d*L -> softmax over L.
    The feature columns of the first layer give one vector per cloud; the three xyz columns are added per point."""
    B, S, _ = sampled_xyz.shape
    m0, m2, m4 = model_mlp[0], model_mlp[2], model_mlp[4]
    w0 = _w2d(m0)
    per_cloud = mlp_ops.linear_small(feature, w0[:, 3:], m0.bias)                   # [B, 512]
    h = mlp_ops.fold_first(sampled_xyz.reshape(B * S, 3), w0[:, :3], per_cloud, S, relu=True)
    h = mlp_ops.linear(h, _w2d(m2), m2.bias, True)
    logits = mlp_ops.linear(h, _w2d(m4), m4.bias, False, out_f32=True)             # fp32 logits [B*S, d*L]
    return torch.softmax(logits.reshape(B, S, d, L), dim=3)


def prob_forward(mod, sampled_xyz):
    """AE.ConditionalProbabilityModel.forward (AE.py:107-123): sampled_xyz [B, S, 3] -> pmf [B, S, d, L].
    Every kernel on this path computes a row (or a cloud) independently of how many there are in the call, in a fixed
    order: the PMFs -- and so the coded stream -- do not depend on the batch size."""
    B, S, _ = sampled_xyz.shape
    xyz = sampled_xyz.detach().float().contiguous()
    feature = pointnet_points(mod.model_pn, xyz)                                    # AE.py:112  [B, 256]
    return _prob_tail(mod.model_mlp, xyz, feature, mod.d, mod.L)


def pppf_prob_forward(mod, sampled_xyz):
    """PPPF_AE.ConditionalProbabilityModel.forward (PPPF_AE.py:203-228)."""
    xyz = sampled_xyz.detach().float().contiguous()
    _, feature = pointnetpp_forward(mod.model_pnpp, xyz)
    return _prob_tail(mod.model_mlp, xyz, feature.contiguous(), mod.d, mod.L)


# ---- pointnet_sa_module.PointnetSAModule, PPPF_AE ---------------------------------------------------------------------------
def sa_module_forward(mod, xyz, features=None):
    """PointnetSAModule.forward (pointnet_sa_module.py:58-93): xyz [B,N,3], features [B,C,N] or None ->
    (new_xyz [B,npoint,3], new_features [B,C_out,npoint])."""
    xyz = xyz.detach()
    B, N, _ = xyz.shape
    fps_idx, new_xyz = ops.fps(xyz, mod.npoint, None, ops.FLT_MAX, return_xyz=True)     # :66-68 (start index 0)
    if mod.npoint > N:                                                                  # :67 clamp: pads read point 0
        new_xyz = ops.gather(xyz, fps_idx.clamp(min=0))
    _, idx = ops.ball_query(new_xyz, xyz, mod.nsample, mod.radius, return_dists=False)  # :71
    idx = idx.clamp(min=0)                                                              # :27 (pads -> point 0)
    layers = triple_layers(mod)
    cin = (features.shape[1] if features is not None else 0) + (3 if mod.use_xyz else 0)
    if cin >= 64:
        # wide stack: one grouping pass (gather + cat + bf16, zero padded to the GEMM's K granule), then every layer on the
        # streamed tensor-core GEMM with the max over nsample fused into the last one           :73-91
        a = mlp_ops.gather_concat_bf16(features.detach().permute(0, 2, 1) if features is not None else None,
                                       xyz if mod.use_xyz else None, idx, (cin + 63) // 64 * 64)
        out = mlp_ops.stream_chain(a, layers, group=mod.nsample)
    else:
        segs = []
        if features is not None:
            segs.append((ops.gather(features.detach().permute(0, 2, 1).contiguous(), idx).view(-1, features.shape[1]), 1))  # :74-77
        if mod.use_xyz:
            segs.append((ops.gather(xyz, idx).view(-1, 3), 1))                          # :80-85 (not recentred)
        out = mlp_ops.run_chain(segs, layers, group=mod.nsample)                        # :89-91
    return new_xyz, out.view(B, mod.npoint, -1).permute(0, 2, 1)


def pointnetpp_forward(mod, xyz, features=None):
    """PPPF_AE.PointNetPP.forward (PPPF_AE.py:39-46)."""
    for sa in (mod.sa1, mod.sa2, mod.sa3):
        xyz, features = sa_module_forward(sa, xyz, features)
    return xyz, torch.max(features, dim=2)[0]


def _folding_grid(mod, batch, device):
    """PPPF_AE.py:80-89; the grid is a constant: built once per (batch, device) and kept on the device."""
    key = (batch, str(device), mod.grid_size)
    hit = mod.__dict__.get("_pcc_grid")
    if hit is None or hit[0] != key:
        x = torch.linspace(-1, 1, mod.grid_size)
        gx, gy = torch.meshgrid(x, x, indexing="ij")
        grid = torch.stack([gx, gy], dim=-1).reshape(-1, 2).unsqueeze(0).repeat(batch, 1, 1).to(device).contiguous()
        hit = mod.__dict__["_pcc_grid"] = (key, grid)
    return hit[1]


def _folding_stage(mlp, local, latent, n_pts):
    """One folding stage (PPPF_AE.py:100-104 / 106-109): Conv1d over cat([local (2 or 3 per point), latent tiled]) -> ReLU ->
    Conv1d -> ReLU -> Conv1d(3).  The latent columns of the first layer are one fp32 vector per cloud; the per-point columns
    are added in fp32 and the sum is emitted as the bf16 operand of the second layer's tensor-core GEMM."""
    n_local = local.shape[1]
    w0 = mlp[0].weight.squeeze(-1)
    per_cloud = mlp_ops.linear_small(latent, w0[:, n_local:], mlp[0].bias)
    h = mlp_ops.fold_first(local, w0[:, :n_local], per_cloud, n_pts, relu=True)
    return mlp_ops.run_chain(h, [(mlp[2].weight.squeeze(-1), mlp[2].bias, True), (mlp[4].weight.squeeze(-1), mlp[4].bias, False)])


def folding_forward(mod, latent):
    """PPPF_AE.FoldingNet.forward (PPPF_AE.py:91-109): latent [B, F] -> [B, grid_size^2, 3]."""
    latent = latent.detach().float().contiguous()
    B = latent.shape[0]
    n_pts = mod.grid_size * mod.grid_size
    grid = _folding_grid(mod, B, latent.device)
    coarse = _folding_stage(mod.mlp1, grid.view(B * n_pts, 2), latent, n_pts)       # [B*N, 3] fp32
    fine = _folding_stage(mod.mlp2, coarse, latent, n_pts)
    return fine.view(B, n_pts, 3)


def pppf_forward(mod, xyz):
    """PPPF_AE.PPPF_AE.forward (PPPF_AE.py:128-150): (recon, latent, latent_quantized)."""
    _, latent = pointnetpp_forward(mod.encoder, xyz)
    spread = mod.L - 0.2
    latent = torch.sigmoid(latent) * spread - spread / 2                             # :136-137
    z = mlp_ops.linear_small(latent, mod.enc_proj.weight, mod.enc_proj.bias)        # :139
    latent_q = z.round()                                                            # :142
    dec = mlp_ops.linear_small(latent_q, mod.dec_proj.weight, mod.dec_proj.bias)    # :145
    return folding_forward(mod.decoder, dec), latent, latent_q


# ---- pppe_pcd_ae.PointNet2EncoderFull ("fast pppe_pcd_ae compress", BASELINE cfg5) -----------------------------------------------
def pppe_sa_layers(mod):
    """pppe_pcd_ae.PointNetSetAbstraction.mlp_stack: Sequential(Conv2d(bias=False), BatchNorm2d, ReLU) per layer, or
    Sequential(Conv2d, ReLU) with bn=False (pppe_pcd_ae.py:556-561, 582-586); eval-mode BatchNorm folded into the conv."""
    return _cached(mod, "pppe_stack", lambda: [_seq_layer(s) for s in mod.mlp_stack])


def pppe_sa_points(mod, xyz, feat=None, centres_idx=None):
    """pppe_pcd_ae.PointNetSetAbstraction.forward (pppe_pcd_ae.py:588-618), channel-last: xyz [B, N, 3], feat [B, N, C] or None
    -> (new_xyz [B, S, 3], new features [B, S, C_out] fp32).  FPS start from the CPU RNG (pn_kit.py:321), kNN(K) of every centre,
    recentred xyz [| gathered features], shared MLP, max over the K neighbours.  `centres_idx` [B, S]: the sampling, when the
    caller has already run it (the MSG level batches its branches' samplings)."""
    B, N, _ = xyz.shape
    S, K = mod.npoint, mod.K
    if S == N:
        new_xyz = xyz
    else:
        new_xyz = ops.gather(xyz, centres_idx if centres_idx is not None else pn_kit_ops.farthest_point_sample_batch(xyz, S))   # :596-600
    layers = list(pppe_sa_layers(mod))
    if feat is None:
        _, _, grouped = ops.knn(new_xyz, xyz, K, return_nn=True, centre_sub=True, nn_only=True)       # :602-603
        out = mlp_ops.run_chain(grouped.reshape(B * S * K, 3), layers, group=K)                      # :614-617
    else:
        _, idx, _ = ops.knn(new_xyz, xyz, K)                                                          # :602
        C = feat.shape[2]

        def rot():   # the reference concatenates [xyz | features] (:609); the grouping kernel emits [features | xyz]
            w0 = layers[0][0]
            return torch.cat((w0[:, 3:], w0[:, :3]), dim=1).detach().contiguous()

        layers[0] = (_cached(mod, "pppe_rot_w0", rot), layers[0][1], layers[0][2])
        a = mlp_ops.gather_concat_bf16(feat, xyz, idx, (C + 3 + 63) // 64 * 64, centre=new_xyz, nsample=K)   # :603-609
        out = mlp_ops.stream_chain(a, layers, group=K)                                                # :614-617
    return new_xyz, out.reshape(B, S, -1)


# Above this many points a cloud's sampling occupies most of the GPU (co-resident CTAs, 8192 points each): samplings of the same
# cloud are then cheaper as ONE batched call, which the scene-scale form runs side by side on one SM each (fps_bucket.cu).
_MSG_BATCHED_FPS_MIN_POINTS = 196608


def pppe_msg_points(mod, xyz, feat=None):
    """pppe_pcd_ae.PointNetSetAbstractionMSG.forward (pppe_pcd_ae.py:627-633): every branch samples its own centres (one CPU-RNG
    draw each, in branch order); the LAST branch's centres are handed on; features concatenated along the channel axis.  On
    scene-sized clouds the branches' samplings -- independent sequences over the same points -- run as one batched FPS call."""
    B, N, _ = xyz.shape
    branches = list(mod.branches)
    samplings = [None] * len(branches)
    nps = {b.npoint for b in branches}
    if len(branches) > 1 and len(nps) == 1 and N >= _MSG_BATCHED_FPS_MIN_POINTS and branches[0].npoint != N:
        starts = torch.cat([torch.randint(0, N, (B,), dtype=torch.long) for _ in branches])       # pn_kit.py:321, one draw per branch
        idx = ops.fps(xyz.repeat(len(branches), 1, 1), branches[0].npoint, starts.to(xyz.device), 1e10)
        samplings = list(idx.reshape(len(branches), B, -1).unbind(0))
    outs, new_xyz = [], None
    for b, ci in zip(branches, samplings):
        new_xyz, f = pppe_sa_points(b, xyz, feat, ci)
        outs.append(f)
    return new_xyz, torch.cat(outs, dim=2)


def _pppe_level(mod, xyz, feat):
    return pppe_msg_points(mod, xyz, feat) if hasattr(mod, "branches") else pppe_sa_points(mod, xyz, feat)


def pppe_sa_forward(mod, xyz, points=None):
    """Reference layouts: xyz [B, N, 3], points [B, C, N] or None -> (new_xyz [B, S, 3], new_points [B, C_out, S])."""
    xyz = xyz.detach().float().contiguous()
    feat = points.detach().float().permute(0, 2, 1).contiguous() if points is not None else None
    new_xyz, f = _pppe_level(mod, xyz, feat)
    return new_xyz, f.permute(0, 2, 1)


def pppe_encoder_forward(mod, x):
    """pppe_pcd_ae.PointNet2EncoderFull.forward (pppe_pcd_ae.py:672-690): x [B, N, 3] -> (latent [B, latent_dim], pooled features
    [B, C_out]).  The levels hand channel-last features to each other (no permutes in between)."""
    xyz, feat = x.detach().float().contiguous(), None
    for sa in mod.sa_modules:
        xyz, feat = _pppe_level(sa, xyz, feat)
    global_feat = feat.max(dim=1)[0]                                                                  # :683
    conv0, bn, conv3 = mod.global_conv[0], mod.global_conv[1], mod.global_conv[3]
    w0, b0 = _cached(mod, "pppe_global", lambda: _fold(conv0, bn))
    h = mlp_ops.linear_small(global_feat, w0, b0, relu=True)                                          # :660-665, 685
    latent = mlp_ops.linear_small(h, _w2d(conv3), conv3.bias)
    return latent, global_feat
