"""Training form of the shared-MLP stacks: forward AND backward on the hand-written tensor-core kernels.

The reference trains its Conv2d(1x1) / Linear stacks under autograd (/root/reference/train.py:193-221 through pn_kit.py:124-211,
289-305 and AE.py:19-55).  `mlp_train` is that stack as ONE autograd function whose every contraction is a pcc kernel:

    forward   y_l = act(x_l W_l^T + b_l)            pcc_linear_train_bf16 (csrc/gemm_ws.cu), bf16 activations kept for backward
              max over runs of `group` rows         pcc_groupmax_fwd_bf16 (csrc/train_ops.cu)
    backward  dY_L from the pooled gradient         pcc_groupmax_bwd_bf16
              dW_l = dY_l^T x_l, db_l = colsum dY_l pcc_wgrad_bf16 (csrc/wgrad_ws.cu: MN-major tcgen05 contraction over the rows)
              dY_{l-1} = (dY_l W_l) * [x_l > 0]     pcc_linear_train_bf16 with the ReLU mask fused into its epilogue

Numerics: bf16 operands (activations, gradients, per-step bf16 copies of the fp32 master weights), fp32 accumulation, fp32 weight
gradients -- the counterpart of the reference's fp16 autocast path (train.py:114,154); tolerances are stated in
tests/test_gpu_train.py.  There is no library GEMM in here.
"""
import torch

from . import _lib, mlp_ops


def _ru(a, b):
    return (a + b - 1) // b * b


_zero_bias = {}


def _zeros(n, device):
    key = (n, str(device))
    z = _zero_bias.get(key)
    if z is None:
        z = _zero_bias[key] = torch.zeros((n,), dtype=torch.float32, device=device)
    return z


def linear_train(x, w, b, relu, mask=None, n_store=None):
    """bf16 [M, n_store] = act(x [M, Kp] . w [cout, cin]^T + b) (columns >= cout are zero), optionally zeroed where mask <= 0."""
    lib = _lib.load()
    M, kp = x.shape
    cout, cin = w.shape
    if x.dtype != torch.bfloat16 or x.stride(1) != 1 or kp % 64 or x.stride(0) % 8 or x.data_ptr() % 16 or cin > kp:
        raise ValueError("pcc_b200.linear_train: x must be a 16-byte aligned bf16 [M, K] tensor with K % 64 == 0 and K >= cin")
    npad = _ru(cout, 128)
    n_store = _ru(cout, 64) if n_store is None else n_store
    wp, bp = mlp_ops._w_bf16_padded(w, b if b is not None else _zeros(cout, x.device), kp, npad)
    out = torch.empty((M, n_store), dtype=torch.bfloat16, device=x.device)
    if mask is not None and (mask.dtype != torch.bfloat16 or mask.stride(1) != 1 or mask.shape[0] != M or mask.shape[1] < n_store):
        raise ValueError("pcc_b200.linear_train: mask must be bf16 [M, >= n_store]")
    with torch.cuda.device(x.device):
        _lib.check(lib.pcc_linear_train_bf16(x.data_ptr(), M, kp, x.stride(0), wp.data_ptr(), kp, bp.data_ptr(), npad, int(bool(relu)),
                                             out.data_ptr(), n_store, n_store, mask.data_ptr() if mask is not None else None,
                                             mask.stride(0) if mask is not None else 0, torch.cuda.current_stream().cuda_stream),
                   "pcc_linear_train_bf16")
    return out


def groupmax_fwd(x, C, group):
    lib = _lib.load()
    M = x.shape[0]
    pooled = torch.empty((M // group, C), dtype=torch.float32, device=x.device)
    arg = torch.empty((M // group, C), dtype=torch.int16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.pcc_groupmax_fwd_bf16(x.data_ptr(), M, C, x.stride(0), group, pooled.data_ptr(), arg.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream), "pcc_groupmax_fwd_bf16")
    return pooled, arg


def groupmax_bwd(dout, pooled, arg, M, C, group, ld):
    lib = _lib.load()
    dy = torch.empty((M, ld), dtype=torch.bfloat16, device=dout.device)
    dout = dout.float().contiguous()
    with torch.cuda.device(dout.device):
        _lib.check(lib.pcc_groupmax_bwd_bf16(dout.data_ptr(), pooled.data_ptr() if pooled is not None else None, arg.data_ptr(), M, C,
                                             group, dy.data_ptr(), ld, torch.cuda.current_stream().cuda_stream), "pcc_groupmax_bwd_bf16")
    return dy


class _MlpTrain(torch.autograd.Function):
    """x0 bf16 [M, K0p] -> the stack's output; see mlp_train."""

    @staticmethod
    def forward(ctx, x0, group, relus, mode, x0_is_relu, *params):
        L = len(relus)
        acts = [x0]
        for l in range(L - 1):
            acts.append(linear_train(acts[-1], params[2 * l], params[2 * l + 1], relus[l]))
        w, b = params[2 * (L - 1)], params[2 * (L - 1) + 1]
        cout = w.shape[0]
        ctx.group, ctx.relus, ctx.mode, ctx.M, ctx.x0_is_relu = group, relus, mode, x0.shape[0], x0_is_relu
        extra = []
        if mode == "pool":
            if cout % 8:
                raise ValueError("pcc_b200.mlp_train: a pooled stack needs an output width that is a multiple of 8")
            last = linear_train(acts[-1], w, b, relus[-1])
            out, arg = groupmax_fwd(last, cout, group)
            extra = [arg]
        elif mode == "bf16":
            out = linear_train(acts[-1], w, b, relus[-1])
        else:   # "f32": values that must not be rounded (decoded coordinates, logits)
            out = mlp_ops.linear(acts[-1], w, b if b is not None else _zeros(cout, x0.device), relus[-1], out_f32=True).contiguous()
        ctx.has_bias = tuple(p is not None for p in params)
        ctx.n_acts, ctx.n_extra = len(acts), len(extra)
        # everything through save_for_backward (the output included: an attribute would make a reference cycle that keeps the
        # parameters' AccumulateGrad nodes alive across steps and breaks CUDA-graph capture of the training step)
        ctx.save_for_backward(*acts, *extra, out, *[p for p in params if p is not None])
        return out

    @staticmethod
    def backward(ctx, dout):
        saved = ctx.saved_tensors
        acts = list(saved[:ctx.n_acts])
        extra = saved[ctx.n_acts:ctx.n_acts + ctx.n_extra]
        out = saved[ctx.n_acts + ctx.n_extra]
        it = iter(saved[ctx.n_acts + ctx.n_extra + 1:])
        params = [next(it) if h else None for h in ctx.has_bias]
        relus, L = ctx.relus, len(ctx.relus)
        cout = params[2 * (L - 1)].shape[0]
        ld = _ru(cout, 64)
        if ctx.mode == "pool":
            dy = groupmax_bwd(dout, out if relus[-1] else None, extra[0], ctx.M, cout, ctx.group, ld)
        else:
            g = dout
            if relus[-1]:                                  # ReLU on the last layer
                g = g * (out[:, :g.shape[1]] > 0)
            if g.dtype == torch.bfloat16 and g.shape[1] == ld and g.is_contiguous():
                dy = g
            else:
                dy = torch.zeros((ctx.M, ld), dtype=torch.bfloat16, device=dout.device)
                dy[:, :cout] = g[:, :cout]
        grads = [None] * (2 * L)
        dx0 = None
        for l in range(L - 1, -1, -1):
            w, b = params[2 * l], params[2 * l + 1]
            co, ci = w.shape
            if ctx.needs_input_grad[5 + 2 * l] or (b is not None and ctx.needs_input_grad[6 + 2 * l]):
                dw, db = mlp_ops.wgrad(dy[:, :_ru(co, 8)], acts[l], want_bias=b is not None)
                grads[2 * l] = dw[:co, :ci].to(w.dtype)
                if b is not None:
                    grads[2 * l + 1] = db[:co].to(b.dtype)
            if l > 0:
                dy = linear_train(dy, w.t(), None, False, mask=acts[l] if relus[l - 1] else None, n_store=acts[l].shape[1])
            elif ctx.needs_input_grad[0]:
                dx0 = linear_train(dy, w.t(), None, False, mask=acts[0] if ctx.x0_is_relu else None, n_store=acts[0].shape[1])
        return (dx0, None, None, None, None) + tuple(grads)


def mlp_train(x0, layers, group=0, mode="f32", x0_is_relu=False):
    """Differentiable shared-MLP stack on the pcc kernels.

    x0      bf16 [M, K0p]: the first layer's input, zero padded to a multiple of 64 columns (gradient: bf16, same shape);
    layers  [(weight [cout, cin], bias [cout] or None, relu)], fp32 parameters (cin_0 <= K0p; later layers read the previous
            layer's channels);
    group / mode   "pool": max over every run of `group` rows -> fp32 [M / group, cout];  "bf16": bf16 [M, roundup(cout, 64)]
            (feeds another stack);  "f32": fp32 [M, cout];
    x0_is_relu     x0 is itself the output of a ReLU layer (fold_first_train): its gradient is masked where x0 <= 0."""
    if mode == "pool" and group <= 1:
        raise ValueError("pcc_b200.mlp_train: mode 'pool' needs group > 1")
    flat = []
    for w, b, _ in layers:
        flat += [w, b]
    return _MlpTrain.apply(x0, int(group), tuple(bool(r) for _, _, r in layers), mode, bool(x0_is_relu), *flat)


class _FoldFirst(torch.autograd.Function):
    """relu(local [M, n <= 4] fp32 . w [C, n]^T + b) -> bf16 [M, roundup(C, 64)] on the CUDA cores (csrc/small_ops.cu: the
    coordinates are not rounded to bf16, as in the inference chain's first layer); backward: the incoming gradient is already
    masked by the consumer (mlp_train(..., x0_is_relu=True)), dW / db come from the weight-gradient kernel."""

    @staticmethod
    def forward(ctx, local, w, b):
        M = local.shape[0]
        out = mlp_ops.fold_first(local, w, b.detach().float().reshape(1, -1), M, relu=True)
        ctx.save_for_backward(torch.nn.functional.pad(local.detach().to(torch.bfloat16), (0, 8 - local.shape[1])))
        ctx.shape = tuple(w.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        co, ci = ctx.shape
        dw, db = mlp_ops.wgrad(dout[:, :_ru(co, 8)], ctx.saved_tensors[0])
        return None, dw[:co, :ci], db[:co]


def fold_first_train(local, w, b):
    """First layer of a stack whose input is a few fp32 coordinates per row (SetAbstraction's recentred neighbours)."""
    return _FoldFirst.apply(local.contiguous(), w, b)


class _SaIndexed(torch.autograd.Function):
    """pn_kit.SetAbstraction's shared MLP + max over the neighbours as one forward kernel (mlp_ops.sa_chain_indexed, the inference
    kernel: nothing is kept but the inputs) and one backward kernel that recomputes the activations (csrc/sa_bwd.cu)."""

    @staticmethod
    def forward(ctx, patches, idx8, w0, b0, w1, b1, w2, b2, out_bf16):
        layers = [(w0.detach(), b0.detach(), True), (w1.detach(), b1.detach(), True), (w2.detach(), b2.detach(), True)]
        out = mlp_ops.sa_chain_indexed(patches, idx8, layers, out_dtype=torch.bfloat16 if out_bf16 else torch.float32)
        ctx.save_for_backward(patches, idx8, w0, b0, w1, b1, w2, b2)
        return out

    @staticmethod
    def backward(ctx, dout):
        patches, idx8, *params = ctx.saved_tensors
        grads = mlp_ops.sa_chain_indexed_bwd(patches, idx8, params, dout)
        return (None, None) + tuple(g.to(p.dtype).reshape(p.shape) for g, p in zip(grads, params)) + (None,)


def sa_indexed_train(patches, idx8, layers, out_bf16=False):
    """Differentiable SetAbstraction stack (3 -> 32 -> 64 -> 128, ReLU everywhere, max over the 16 in-patch neighbours of idx8):
    patches [BS, P, 3] fp32 (no gradient), layers [(weight, bias, True)] x 3 -> [BS * P, 128], fp32 or -- out_bf16: the kernel rounds
    its fp32 result once, exactly what a `.to(torch.bfloat16)` of the fp32 output would hold, without the 268 MB fp32 round trip."""
    (w0, b0, _), (w1, b1, _), (w2, b2, _) = layers
    return _SaIndexed.apply(patches, idx8, w0, b0, w1, b1, w2, b2, bool(out_bf16))


def pad_bf16(t, width):
    """fp32 / bf16 [M, C] -> bf16 [M, width] with zero columns past C (differentiable)."""
    t = t.to(torch.bfloat16)
    return torch.nn.functional.pad(t, (0, width - t.shape[1])) if t.shape[1] < width else t
