"""Shared-MLP chains (1x1-conv stacks) on the device: the dense contractions of the hot path.

    run_chain(inputs, layers, group)              -> [M, CL] or [M / group, CL]  (max over runs of `group` consecutive rows)
    mlp_chain / mlp_chain_groupmax                   the same for one [M, C0] input tensor

`layers` is a list of (weight [Cout, Cin], bias [Cout], relu: bool) -- the reference's Conv2d / Conv1d / Linear parameters.

Three tcgen05 kernels cover every layer, all hand-written (bf16 operands, fp32 accumulation in TMEM):
  * fused_chain  -- runs of layers whose packed weights fit in shared memory, ONE launch, activations never leave the SM
                    (csrc/mlp_chain.cu; the AE's three chains dispatch to the warp-specialised kernels of csrc/chain_ws.cu);
  * pn_tail      -- PointNet's 256 -> 512 -> d + max over the patch (csrc/pn_tail.cu), W2 streamed through a TMA ring;
  * linear       -- one wide layer as a streamed GEMM (csrc/gemm_ws.cu): weights too large to sit beside the activations
                    (AE.inv_pool's 1024 -> 16384, PointNet++'s 128..1024-wide layers, FoldingNet's 512 -> 512).
Skinny per-cloud vectors (one row per cloud) and the per-point part of a "cat(point, tiled cloud vector)" first layer run in
fp32 on the CUDA cores (linear_small, fold_first: csrc/small_ops.cu).
There is no library GEMM and no CPU path: a layer none of the kernels takes raises ValueError (PCC_ERR_UNSUPPORTED).
"""
import weakref

import torch

from . import _lib

_P = 128
_SMEM_MAX = 227 * 1024
_CHUNK_ROWS = 1 << 20
_pack_cache = {}


def _check(x):
    if not x.is_cuda:
        raise RuntimeError("pcc_b200: mlp_chain input must be a CUDA tensor (there is no CPU path)")


def _ru(a, b):
    return (a + b - 1) // b * b


def _fused_smem_bytes(dims, pooled=False):
    """Shared memory the fused kernel needs for a chain with channel sizes dims = [C0, C1, ..., CL]
    (mirrors pcc_mlp_chain: resident packed weights + one activation buffer)."""
    n = len(dims) - 1
    w = 0
    for l, (ci, co) in enumerate(zip(dims[:-1], dims[1:])):
        rows = _ru(co, 128) if (pooled and l == n - 1) else _ru(co, 16)
        w += rows * _ru(ci + 1, 16) * 2
    x0 = _P * _ru(dims[0] + 1, 16) * 2
    x = _P * max([_ru(c + 1, 16) for c in dims[1:-1]] or [0]) * 2
    return _ru(w, 128) + x0 + x + 32 + 4096  # + first-layer fp32 weights when that layer runs on the CUDA cores


def _fits(dims, pooled=False):
    return (len(dims) - 1 <= 6 and _fused_smem_bytes(dims, pooled) <= _SMEM_MAX and max(dims[1:]) <= 512)


def _base(t):
    return t._base if t._base is not None else t


def _packed(w, b):
    """Pack (and cache) one layer's weights and bias for the tcgen05 kernel.  The cache is keyed on the parameter
    tensors that own the storage (`w` is usually a fresh `.flatten(1)` view of a Conv2d weight on every call) plus the
    view geometry, and holds weak references + version counters, so an in-place update or a recycled id re-packs."""
    bw, bb = _base(w), _base(b)
    key = (id(bw), id(bb), w.data_ptr(), tuple(w.shape), tuple(w.stride()), b.data_ptr())
    hit = _pack_cache.get(key)
    if hit is not None:
        rw, rb, vw, vb, val = hit
        if rw() is bw and rb() is bb and vw == bw._version and vb == bb._version:
            return val
    lib = _lib.load()
    cout, cin = w.shape
    wf = w.detach().float().contiguous()
    bf = b.detach().float().contiguous()
    buf = torch.empty((lib.pcc_mlp_packed_bytes(cin, cout),), dtype=torch.uint8, device=w.device)
    with torch.cuda.device(w.device):
        _lib.check(lib.pcc_mlp_pack_weights_f32(wf.data_ptr(), bf.data_ptr(), cin, cout, buf.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream), "pcc_mlp_pack_weights_f32")
    if len(_pack_cache) > 256:
        _pack_cache.clear()
    val = (buf, wf, bf)
    _pack_cache[key] = (weakref.ref(bw), weakref.ref(bb), bw._version, bb._version, val)
    return val


def fused_chain(inputs, layers, group=0, out_dtype=torch.float32):
    """One launch of the tcgen05 chain kernel.

    inputs: a [M, C0] tensor, or a list of segments `(tensor [R, C], row_div)` concatenated along the channel axis
    (segment row used for position r is r // row_div); fp32 or bf16, unit stride along channels.
    Returns [M, CL] (group <= 1) or [M / group, CL], fp32 or bf16."""
    lib = _lib.load()
    if isinstance(inputs, torch.Tensor):
        inputs = [(inputs, 1)]
    segs = (_lib.PccMlpInput * len(inputs))()
    keep = []
    M = None
    for i, (t, div) in enumerate(inputs):
        _check(t)
        if t.dtype not in (torch.float32, torch.bfloat16):
            t = t.float()
        if t.dim() != 2 or t.stride(1) != 1:
            t = t.reshape(-1, t.shape[-1]).contiguous()
        keep.append(t)
        rows_here = t.shape[0] * div
        M = rows_here if M is None else M
        if rows_here != M:
            raise ValueError("pcc_b200.fused_chain: input segments disagree on the number of rows")
        segs[i] = _lib.PccMlpInput(t.data_ptr(), 0 if t.dtype == torch.float32 else 1, t.shape[1], t.stride(0), div)
    arr = (_lib.PccMlpLayer * len(layers))()
    for i, (w, b, relu) in enumerate(layers):
        pw, wf, bf = _packed(w, b)
        keep.append((pw, wf, bf))
        arr[i] = _lib.PccMlpLayer(pw.data_ptr(), w.shape[1], w.shape[0], int(bool(relu)), wf.data_ptr(), bf.data_ptr())
    cl = layers[-1][0].shape[0]
    out_rows = M // group if group > 1 else M
    dev = keep[0].device
    out = torch.empty((out_rows, cl), dtype=out_dtype, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.pcc_mlp_chain(segs, len(inputs), M, arr, len(layers), int(group), out.data_ptr(),
                                     0 if out_dtype == torch.float32 else 1, torch.cuda.current_stream().cuda_stream),
                   "pcc_mlp_chain")
    return out


def sa_indexed_supported(P, K, layers):
    """Shapes pcc_sa_chain_indexed takes: pn_kit.SetAbstraction as the AE uses it (3 -> 32 -> 64 -> 128, ReLU, K = 16)."""
    dims = [layers[0][0].shape[1]] + [w.shape[0] for w, _, _ in layers]
    return (K == 16 and 16 <= P <= 256 and P % 8 == 0 and dims == [3, 32, 64, 128] and all(bool(r) for _, _, r in layers))


def sa_chain_indexed(patches, idx8, layers, out_dtype=torch.float32):
    """SetAbstraction shared MLP + max over the 16 neighbours, gathering the recentred neighbours from the patch itself:
    patches [BS, P, 3] fp32, idx8 [BS, P, 16] uint8 (ops.knn_patch_u8) -> [BS * P, 128]."""
    lib = _lib.load()
    _check(patches)
    BS, P, _ = patches.shape
    patches = patches.float().contiguous()
    idx8 = idx8.contiguous()
    arr = (_lib.PccMlpLayer * len(layers))()
    keep = []
    for i, (w, b, relu) in enumerate(layers):
        pw, wf, bf = _packed(w, b)
        keep.append((pw, wf, bf))
        arr[i] = _lib.PccMlpLayer(pw.data_ptr(), w.shape[1], w.shape[0], int(bool(relu)), wf.data_ptr(), bf.data_ptr())
    out = torch.empty((BS * P, layers[-1][0].shape[0]), dtype=out_dtype, device=patches.device)
    with torch.cuda.device(patches.device):
        _lib.check(lib.pcc_sa_chain_indexed(patches.data_ptr(), idx8.data_ptr(), BS * P, P, arr, len(layers), out.data_ptr(),
                                            0 if out_dtype == torch.float32 else 1, torch.cuda.current_stream().cuda_stream),
                   "pcc_sa_chain_indexed")
    return out


def sa_chain_indexed_bwd(patches, idx8, params, grad_out):
    """Parameter gradients of sa_chain_indexed (csrc/sa_bwd.cu): params = (w0 [32,3], b0, w1 [64,32], b1, w2 [128,64], b2) fp32,
    grad_out [BS * P, 128] -> six fp32 tensors shaped like params."""
    lib = _lib.load()
    _check(patches)
    BS, P, _ = patches.shape
    patches = patches.float().contiguous()
    idx8 = idx8.contiguous()
    ps = [t.detach().float().contiguous() for t in params]
    if [tuple(t.shape) for t in ps] != [(32, 3), (32,), (64, 32), (64,), (128, 64), (128,)]:
        raise ValueError("pcc_b200.sa_chain_indexed_bwd: the stack must be 3 -> 32 -> 64 -> 128")
    g = grad_out.detach()
    if tuple(g.shape) != (BS * P, 128):
        raise ValueError("pcc_b200.sa_chain_indexed_bwd: grad_out must be [BS * P, 128]")
    if g.dtype not in (torch.float32, torch.bfloat16) or g.stride(1) != 1 or g.stride(0) < 128:
        g = g.float().contiguous()   # fp32 or bf16 rows are read in place (a column slice of a wider gradient included)
    grads = [torch.zeros_like(t) for t in ps]
    with torch.cuda.device(patches.device):
        _lib.check(lib.pcc_sa_chain_indexed_bwd(patches.data_ptr(), idx8.data_ptr(), BS * P, P, *[t.data_ptr() for t in ps], g.data_ptr(),
                                                0 if g.dtype == torch.float32 else 1, g.stride(0),
                                                *[t.data_ptr() for t in grads], torch.cuda.current_stream().cuda_stream),
                   "pcc_sa_chain_indexed_bwd")
    return grads


_bf16_cache = {}


def _bf16(t):
    bt = _base(t)
    key = (id(bt), t.data_ptr(), tuple(t.shape), tuple(t.stride()))
    hit = _bf16_cache.get(key)
    if hit is not None and hit[0]() is bt and hit[1] == bt._version:
        return hit[2]
    if len(_bf16_cache) > 256:
        _bf16_cache.clear()
    val = t.detach().to(torch.bfloat16).contiguous()
    _bf16_cache[key] = (weakref.ref(bt), bt._version, val)
    return val


def pn_tail_supported(x, layers, group):
    """The fused PointNet tail kernel: [M, 256] bf16 -> 512 (ReLU) -> cout <= 16, max over runs of 256 rows."""
    return (len(layers) == 2 and group == 256 and x.dtype == torch.bfloat16 and x.dim() == 2 and x.shape[1] == 256 and
            x.stride(1) == 1 and x.stride(0) % 8 == 0 and x.data_ptr() % 16 == 0 and x.shape[0] % 256 == 0 and
            tuple(layers[0][0].shape) == (512, 256) and layers[0][2] and layers[1][0].shape[1] == 512 and
            layers[1][0].shape[0] <= 16)


def pn_tail(x, layers):
    """relu(W2 x + b2) -> W3 . + b3 -> max over each run of 256 rows, in one launch (csrc/pn_tail.cu)."""
    lib = _lib.load()
    (w2, b2, _), (w3, b3, relu3) = layers
    w2h = _bf16(w2)
    b2f = b2.detach().float().contiguous()
    pw3, _, _ = _packed(w3, b3)
    out = torch.empty((x.shape[0] // 256, w3.shape[0]), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.pcc_pn_tail_bf16(x.data_ptr(), x.shape[0], x.stride(0), w2h.data_ptr(), b2f.data_ptr(), pw3.data_ptr(),
                                        w3.shape[0], int(bool(relu3)), out.data_ptr(), torch.cuda.current_stream().cuda_stream),
                   "pcc_pn_tail_bf16")
    return out


def pointnet_fused_supported(feat, xyz, layers, group):
    """The one-launch PointNet of the AE: [feat 128 bf16 | xyz 3] -> 128 -> 256 -> 512 -> cout <= 16, ReLU on the first three,
    max over runs of 256 rows."""
    dims = [layers[0][0].shape[1]] + [w.shape[0] for w, _, _ in layers]
    return (len(layers) == 4 and group == 256 and dims[:4] == [131, 128, 256, 512] and dims[4] <= 16 and
            all(bool(r) for _, _, r in layers[:3]) and feat.dtype == torch.bfloat16 and feat.dim() == 2 and feat.shape[1] == 128 and
            feat.stride(1) == 1 and feat.stride(0) % 8 == 0 and feat.data_ptr() % 16 == 0 and feat.shape[0] % 256 == 0 and
            xyz.shape[0] == feat.shape[0] and xyz.shape[1] == 3)


def pointnet_fused(feat, xyz, layers):
    """pn_kit.PointNet on [feat | xyz] in one launch (csrc/pn_fused.cu); layers[0]'s weight columns are [feat (128) | xyz (3)]."""
    lib = _lib.load()
    (w0, b0, _), (w1, b1, _), (w2, b2, _), (w3, b3, relu3) = layers
    w0f = _bf16(w0[:, :128])
    pw0, _, _ = _packed(w0, b0)
    w1h, w2h = _bf16(w1), _bf16(w2)
    b1f, b2f = b1.detach().float().contiguous(), b2.detach().float().contiguous()
    pw3, _, _ = _packed(w3, b3)
    xyz = xyz.detach().float()
    if xyz.stride(1) != 1:
        xyz = xyz.contiguous()
    out = torch.empty((feat.shape[0] // 256, w3.shape[0]), dtype=torch.float32, device=feat.device)
    with torch.cuda.device(feat.device):
        _lib.check(lib.pcc_pointnet_fused_bf16(feat.data_ptr(), feat.shape[0], feat.stride(0), xyz.data_ptr(), xyz.stride(0),
                                               w0f.data_ptr(), pw0.data_ptr(), w1h.data_ptr(), b1f.data_ptr(), w2h.data_ptr(),
                                               b2f.data_ptr(), pw3.data_ptr(), w3.shape[0], int(bool(relu3)), out.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream), "pcc_pointnet_fused_bf16")
    return out


def _split(layers, pooled=False):
    """Longest prefix of `layers` that fits the fused kernel."""
    dims = [layers[0][0].shape[1]] + [w.shape[0] for w, _, _ in layers]
    n = len(layers)
    while n > 0 and not _fits(dims[:n + 1], pooled and n == len(layers)):
        n -= 1
    return n


def _linear_operand(inputs):
    """The A operand of the streamed GEMM for `inputs` (a tensor or a list of channel segments): 16-byte aligned bf16 [M, Kp]
    with Kp % 64 == 0, zero columns past the data.  A bf16 activation a previous kernel wrote in that form is used as is."""
    if isinstance(inputs, torch.Tensor):
        inputs = [(inputs, 1)]
    if len(inputs) == 1 and inputs[0][1] == 1:
        t = inputs[0][0]
        if (t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1 and t.shape[1] % 64 == 0 and t.stride(0) % 8 == 0 and
                t.data_ptr() % 16 == 0):
            return t
    parts = []
    for t, div in inputs:
        t = t.reshape(-1, t.shape[-1]).to(torch.bfloat16)
        parts.append(t.repeat_interleave(div, dim=0) if div > 1 else t)
    c = sum(p.shape[1] for p in parts)
    out = torch.zeros((parts[0].shape[0], _ru(c, 64)), dtype=torch.bfloat16, device=parts[0].device)
    at = 0
    for p in parts:
        out[:, at:at + p.shape[1]] = p
        at += p.shape[1]
    return out


def _rows(inputs):
    if isinstance(inputs, torch.Tensor):
        return inputs.shape[0]
    t, div = inputs[0]
    return t.reshape(-1, t.shape[-1]).shape[0] * div


def run_chain(inputs, layers, group=0, out_dtype=torch.float32):
    """Shared-MLP chain of any size: greedy runs of layers that fit the fused tcgen05 kernel (activations stay on the SM
    inside a run, bf16 in HBM between runs); a layer whose weights do not fit beside the activations runs on the streamed
    GEMM; PointNet's 256 -> 512 -> d + max tail on its own kernel.  The pooling is always fused into the last launch."""
    L = len(layers)
    cur, i = inputs, 0
    while i < L:
        rest = layers[i:]
        if i > 0 and group > 1 and isinstance(cur, torch.Tensor) and out_dtype == torch.float32 and pn_tail_supported(cur, rest, group):
            return pn_tail(cur, rest)
        n = _split(rest, pooled=group > 1)
        if n > 0:
            last = i + n == L
            cur = fused_chain(cur, rest[:n], group if last else 0, out_dtype if last else torch.bfloat16)
            i += n
        else:
            last = i + 1 == L
            w, b, relu = rest[0]
            cur = linear(_linear_operand(cur), w, b, relu, group if (last and group > 1) else 0,
                         out_f32=last and group <= 1 and out_dtype == torch.float32)
            if last and cur.dtype != out_dtype:
                cur = cur.to(out_dtype)
            i += 1
    return cur


def mlp_chain(x, layers):
    _check(x)
    return run_chain(x, layers, 0)


def mlp_chain_groupmax(x, layers, group):
    _check(x)
    if x.shape[0] % group:
        raise ValueError("pcc_b200.mlp_chain_groupmax: rows must be a multiple of the group size")
    return run_chain(x, layers, group)


# ---- streamed tensor-core GEMM layers (csrc/gemm_ws.cu): weights too large to sit beside the activations ---------------
_wpad_cache = {}
_weights_epoch = 0


def invalidate_weight_caches():
    """Drop every cached bf16 / packed copy of a weight.  The caches are keyed on the parameters' version counters, which an
    in-place update made outside autograd's book-keeping does not bump -- torch.optim.Adam(fused=True) is one: the trainer calls
    this around every optimiser step, and so should anybody who writes into a parameter's storage by other means."""
    global _weights_epoch
    _weights_epoch += 1        # part of bodies._state_key: per-module derived weights and captured graphs are rebuilt too
    _pack_cache.clear()
    _wpad_cache.clear()
    _bf16_cache.clear()


def _w_bf16_padded(w, b, kpad, npad):
    """[cout, cin] parameter (+ bias) -> cached (bf16 [npad, kpad] with zero rows / columns past the data, fp32 bias [npad])."""
    bw, bb = _base(w), _base(b)
    key = (id(bw), id(bb), w.data_ptr(), b.data_ptr(), tuple(w.shape), tuple(w.stride()), kpad, npad)
    hit = _wpad_cache.get(key)
    if hit is not None and hit[0]() is bw and hit[1]() is bb and hit[2] == (bw._version, bb._version):
        return hit[3]
    if len(_wpad_cache) > 256:
        _wpad_cache.clear()
    wp = torch.zeros((npad, kpad), dtype=torch.bfloat16, device=w.device)
    wp[:w.shape[0], :w.shape[1]] = w.detach()
    bp = torch.zeros((npad,), dtype=torch.float32, device=w.device)
    bp[:w.shape[0]] = b.detach().float()
    _wpad_cache[key] = (weakref.ref(bw), weakref.ref(bb), (bw._version, bb._version), (wp, bp))
    return wp, bp


def linear_supported(rows, cout, group=0):
    """Shapes pcc_linear_bf16 takes (operands are zero padded to its K / N granules by `linear`)."""
    if rows < 1 or cout < 1:
        return False
    if group > 1:
        return group % 32 == 0 and (128 % group == 0 or group % 128 == 0) and rows % group == 0
    return True


def linear(x, w, b, relu, group=0, out_f32=False):
    """One layer on the streamed GEMM kernel.  x [M, Kp] bf16 with Kp % 64 == 0 and Kp >= cin (columns past cin must be
    zero or finite: the weight is zero padded).  Returns bf16 [M, cout]; fp32 [M, cout] with out_f32 (logits, values that must
    not be rounded); fp32 [M / group, cout] when group > 1 (max over runs of `group` rows).  cout is padded to the kernel's
    128-column granule with zero weight rows and the padding is sliced off the result."""
    lib = _lib.load()
    _check(x)
    if x.dtype != torch.bfloat16 or x.dim() != 2 or x.stride(1) != 1 or x.shape[1] % 64 or x.stride(0) % 8 or x.data_ptr() % 16:
        raise ValueError("pcc_b200.linear: x must be a 16-byte aligned bf16 [M, K] tensor with K % 64 == 0")
    M, kp = x.shape
    cout, cin = w.shape
    if cin > kp:
        raise ValueError("pcc_b200.linear: the input has fewer columns than the weight")
    if group > 32 and not relu:
        raise ValueError("pcc_b200.linear: pooling over more than 32 rows needs the ReLU")
    if not linear_supported(M, cout, group):
        raise ValueError(f"pcc_b200.linear: rows={M} cout={cout} group={group} is not a shape the streamed GEMM takes")
    npad = _ru(cout, 128)
    wp, bp = _w_bf16_padded(w, b, kp, npad)
    if group > 1:
        out = torch.empty((M // group, npad), dtype=torch.float32, device=x.device)
    elif out_f32:
        out = torch.empty((M, npad), dtype=torch.float32, device=x.device)
        group = 1
    else:
        out = torch.empty((M, npad), dtype=torch.bfloat16, device=x.device)
        group = 0
    with torch.cuda.device(x.device):
        _lib.check(lib.pcc_linear_bf16(x.data_ptr(), M, kp, x.stride(0), wp.data_ptr(), kp, bp.data_ptr(), npad, int(bool(relu)),
                                       int(group), out.data_ptr(), npad, torch.cuda.current_stream().cuda_stream),
                   "pcc_linear_bf16")
    return out if npad == cout else out[:, :cout]


def linear_small(x, w, b, relu=False):
    """Skinny fp32 Linear for per-cloud vectors (csrc/small_ops.cu): act(x [M, K] . w [N, K]^T + b) -> fp32 [M, N].
    `w` may be a column slice of a wider weight (its row pitch is passed through)."""
    lib = _lib.load()
    _check(x)
    x = x.detach().float()
    if x.dim() != 2 or x.stride(1) != 1:
        x = x.reshape(-1, x.shape[-1]).contiguous()
    w = w.detach()
    if w.dtype != torch.float32 or w.stride(1) != 1:
        w = w.float().contiguous()
    bf = b.detach().float().contiguous() if b is not None else None
    M, K = x.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise ValueError("pcc_b200.linear_small: x and w disagree on K")
    out = torch.empty((M, N), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.pcc_linear_small_f32(x.data_ptr(), M, K, x.stride(0), w.data_ptr(), w.stride(0),
                                            bf.data_ptr() if bf is not None else None, N, int(bool(relu)), out.data_ptr(), N,
                                            torch.cuda.current_stream().cuda_stream), "pcc_linear_small_f32")
    return out


def fold_first(local, w_local, per_cloud, n_pts, relu=True):
    """bf16 [M, roundup(C, 64)] = act(per_cloud[r // n_pts] + local[r] . w_local^T): the first layer of a stage fed
    cat([per-point values, tiled per-cloud vector]) (csrc/small_ops.cu).  local [M, n_local <= 4] fp32, w_local [C, n_local]
    (a column slice of the layer's weight), per_cloud [M / n_pts, C] fp32 including the bias."""
    lib = _lib.load()
    _check(local)
    local = local.detach().float()
    if local.dim() != 2 or local.stride(1) != 1:
        local = local.reshape(-1, local.shape[-1]).contiguous()
    w_local = w_local.detach()
    if w_local.dtype != torch.float32 or w_local.stride(1) != 1:
        w_local = w_local.float().contiguous()
    per_cloud = per_cloud.float().contiguous()
    M, n_local = local.shape
    C = w_local.shape[0]
    if per_cloud.shape != (M // n_pts, C) or w_local.shape[1] != n_local:
        raise ValueError("pcc_b200.fold_first: shapes disagree")
    out = torch.empty((M, _ru(C, 64)), dtype=torch.bfloat16, device=local.device)
    with torch.cuda.device(local.device):
        _lib.check(lib.pcc_fold_first_bf16(local.data_ptr(), n_local, local.stride(0), w_local.data_ptr(), w_local.stride(0),
                                           per_cloud.data_ptr(), M, n_pts, C, int(bool(relu)), out.data_ptr(), out.stride(0),
                                           torch.cuda.current_stream().cuda_stream), "pcc_fold_first_bf16")
    return out


def stream_chain(x, layers, group=0):
    """A run of layers on the streamed GEMM kernel, bf16 in HBM between layers, the pooling fused into the last one."""
    for i, (w, b, relu) in enumerate(layers):
        x = linear(x, w, b, relu, group if i + 1 == len(layers) else 0)
    return x


def gather_concat_bf16(feat, xyz, idx, kpad, centre=None, nsample=1):
    """Rows [feat[b, idx] | xyz[b, idx] | 0...] in bf16, kpad columns (pointnet_sa_module.py:73-85 grouping + cat).
    centre [B, M / nsample, 3]: subtract the query's xyz from the xyz columns (pppe_pcd_ae.py:599-607 recentred grouping)."""
    lib = _lib.load()
    _check(idx)
    B = idx.shape[0]
    idx = idx.reshape(B, -1).contiguous()
    M = idx.shape[1]
    src = feat if feat is not None else xyz
    N = src.shape[1]
    C = 0
    if feat is not None:
        feat = feat.float().contiguous()
        C = feat.shape[2]
    if xyz is not None:
        xyz = xyz.float().contiguous()
    if centre is not None:
        centre = centre.float().contiguous()
    out = torch.empty((B * M, kpad), dtype=torch.bfloat16, device=idx.device)
    with torch.cuda.device(idx.device):
        _lib.check(lib.pcc_gather_concat_bf16(feat.data_ptr() if feat is not None else None, C,
                                              xyz.data_ptr() if xyz is not None else None, idx.data_ptr(), B, N, M, kpad,
                                              out.data_ptr(), centre.data_ptr() if centre is not None else None, int(nsample),
                                              torch.cuda.current_stream().cuda_stream), "pcc_gather_concat_bf16")
    return out


# ---- training: weight gradients (csrc/wgrad_ws.cu) ---------------------------------------------------------------------------
def wgrad(dy, x, want_bias=True):
    """dW [Na, Nb] = dy[M, Na]^T . x[M, Nb] (fp32) and db [Na] = column sums of dy: the weight / bias gradient of a layer
    y = x W^T + b from the bf16 output gradient and the bf16 input activation, both row-major [M, *] (column counts multiples of
    8).  Runs on the tensor cores without transposed copies (MN-major operands)."""
    lib = _lib.load()
    _check(dy)
    for t in (dy, x):
        if t.dtype != torch.bfloat16 or t.dim() != 2 or t.stride(1) != 1 or t.shape[1] % 8 or t.stride(0) % 8 or t.data_ptr() % 16:
            raise ValueError("pcc_b200.wgrad: operands must be 16-byte aligned bf16 [M, C] with C % 8 == 0")
    if dy.shape[0] != x.shape[0]:
        raise ValueError("pcc_b200.wgrad: dy and x disagree on the number of rows")
    M, Na = dy.shape
    Nb = x.shape[1]
    dw = torch.empty((Na, Nb), dtype=torch.float32, device=dy.device)
    db = torch.empty((Na,), dtype=torch.float32, device=dy.device) if want_bias else None
    with torch.cuda.device(dy.device):
        _lib.check(lib.pcc_wgrad_bf16(dy.data_ptr(), dy.stride(0), Na, x.data_ptr(), x.stride(0), Nb, M, dw.data_ptr(), Nb,
                                      db.data_ptr() if db is not None else None, torch.cuda.current_stream().cuda_stream),
                   "pcc_wgrad_bf16")
    return dw, db
