"""Shared-MLP chains (1x1-conv stacks) on the device: the dense contractions of the hot path.

    mlp_chain(x [M, C0], layers)                  -> [M, CL]
    mlp_chain_groupmax(x [M, C0], layers, group)  -> [M/group, CL]   (max over each run of `group` consecutive rows)

`layers` is a list of (weight [Cout, Cin], bias [Cout], relu: bool).

INTERIM (round 1): the contraction itself is issued through torch.addmm on the device (cuBLAS) in row chunks while
the fused tcgen05 kernel is being brought up; the max-pool, layout and chunking live here.  There is still no CPU
path: CPU tensors are rejected.
"""
import torch

_CHUNK_ROWS = 1 << 20


def _check(x):
    if not x.is_cuda:
        raise RuntimeError("pcc_b200: mlp_chain input must be a CUDA tensor (there is no CPU path)")


def _chain(x, layers):
    for w, b, relu in layers:
        x = torch.addmm(b, x, w.t())
        if relu:
            x = torch.relu_(x)
    return x


def mlp_chain(x, layers):
    _check(x)
    if x.shape[0] <= _CHUNK_ROWS:
        return _chain(x, layers)
    return torch.cat([_chain(x[i:i + _CHUNK_ROWS], layers) for i in range(0, x.shape[0], _CHUNK_ROWS)], dim=0)


def mlp_chain_groupmax(x, layers, group):
    _check(x)
    M = x.shape[0]
    if M % group:
        raise ValueError("pcc_b200.mlp_chain_groupmax: rows must be a multiple of the group size")
    step = max(group, (_CHUNK_ROWS // group) * group)
    outs = []
    for i in range(0, M, step):
        y = _chain(x[i:i + step], layers)
        outs.append(y.view(-1, group, y.shape[1]).max(dim=1)[0])
    return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)
