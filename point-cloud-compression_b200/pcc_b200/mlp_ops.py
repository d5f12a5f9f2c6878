"""Shared-MLP chains (1x1-conv stacks) on the device: the dense contractions of the hot path.

    mlp_chain(x [M, C0], layers)                  -> [M, CL]
    mlp_chain_groupmax(x [M, C0], layers, group)  -> [M/group, CL]   (max over each run of `group` consecutive rows)

`layers` is a list of (weight [Cout, Cin], bias [Cout], relu: bool) -- the reference's Conv2d / Linear parameters.

Chains whose packed bf16 weights fit in shared memory run in ONE launch of the fused tcgen05 kernel
(csrc/mlp_chain.cu: bf16 operands, fp32 accumulation in TMEM, activations never leave the SM).  Layers that do not
fit (PointNet's 256->512->16 tail, the 1024->16384 decoder Linear) are still issued as plain library GEMMs
(torch.addmm in bf16 -> cuBLAS) in round 1; the split point is chosen here.  There is no CPU path.
"""
import ctypes
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
    x = _P * max(_ru(c + 1, 16) for c in dims[:-1]) * 2
    return _ru(w, 128) + x + 16


def _fits(dims, pooled=False):
    return (len(dims) - 1 <= 6 and _fused_smem_bytes(dims, pooled) <= _SMEM_MAX and max(dims[1:]) <= 512)


def _packed(w, b):
    """Pack (and cache per parameter tensor + version) one layer's weights and bias for the tcgen05 kernel.  The cache
    entry holds weak references to the tensors, so a recycled address or id can never alias a stale entry."""
    key = (id(w), id(b))
    hit = _pack_cache.get(key)
    if hit is not None:
        rw, rb, vw, vb, buf = hit
        if rw() is w and rb() is b and vw == w._version and vb == b._version:
            return buf, None
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
    _pack_cache[key] = (weakref.ref(w), weakref.ref(b), w._version, b._version, buf)
    return buf, None


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
        pw, pb = _packed(w, b)
        keep.append((pw, pb))
        arr[i] = _lib.PccMlpLayer(pw.data_ptr(), w.shape[1], w.shape[0], int(bool(relu)))
    cl = layers[-1][0].shape[0]
    out_rows = M // group if group > 1 else M
    dev = keep[0].device
    out = torch.empty((out_rows, cl), dtype=out_dtype, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.pcc_mlp_chain(segs, len(inputs), M, arr, len(layers), int(group), out.data_ptr(),
                                     0 if out_dtype == torch.float32 else 1, torch.cuda.current_stream().cuda_stream),
                   "pcc_mlp_chain")
    return out


_bf16_cache = {}


def _bf16(t):
    key = id(t)
    hit = _bf16_cache.get(key)
    if hit is not None and hit[0]() is t and hit[1] == t._version:
        return hit[2]
    if len(_bf16_cache) > 256:
        _bf16_cache.clear()
    val = t.detach().to(torch.bfloat16).contiguous()
    _bf16_cache[key] = (weakref.ref(t), t._version, val)
    return val


def library_chain(x, layers, out_dtype=torch.float32):
    """Plain library GEMMs (cuBLAS through torch) in bf16 with fp32 accumulation, for layers too large for the fused
    kernel's resident-weight design."""
    x = x.to(torch.bfloat16)
    for w, b, relu in layers:
        x = torch.addmm(_bf16(b), x, _bf16(w).t())
        if relu:
            x = torch.relu_(x)
    return x.to(out_dtype)


_library_chain = library_chain


def _split(layers, pooled=False):
    """Longest prefix of `layers` that fits the fused kernel."""
    dims = [layers[0][0].shape[1]] + [w.shape[0] for w, _, _ in layers]
    n = len(layers)
    while n > 0 and not _fits(dims[:n + 1], pooled and n == len(layers)):
        n -= 1
    return n


def mlp_chain(x, layers):
    _check(x)
    n = _split(layers)
    if n == len(layers):
        return fused_chain(x, layers)
    if n > 0:
        x = fused_chain(x, layers[:n])
    if x.shape[0] <= _CHUNK_ROWS:
        return _library_chain(x, layers[n:])
    return torch.cat([_library_chain(x[i:i + _CHUNK_ROWS], layers[n:]) for i in range(0, x.shape[0], _CHUNK_ROWS)])


def mlp_chain_groupmax(x, layers, group):
    _check(x)
    M = x.shape[0]
    if M % group:
        raise ValueError("pcc_b200.mlp_chain_groupmax: rows must be a multiple of the group size")
    n = _split(layers)
    if n == len(layers):
        return fused_chain(x, layers, group)
    if n > 0:
        x = fused_chain(x, layers[:n])
    step = max(group, (_CHUNK_ROWS // group) * group)
    outs = []
    for i in range(0, M, step):
        y = _library_chain(x[i:i + step], layers[n:])
        outs.append(y.view(-1, group, y.shape[1]).max(dim=1)[0])
    return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)
