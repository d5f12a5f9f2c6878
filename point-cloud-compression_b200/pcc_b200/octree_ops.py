"""Drop-ins for the reference's octree centre coder, same names and signatures:

  encode_sampled_np(sampled_xyz, scale, N, min_bpp)   pn_kit.py:380-401   (train.py:176, compress.py:98)
  decode_sampled_np(codes, scale)                     pn_kit.py:424-431   (train.py:177, compress.py:100, decompress.py:83)
  encode(pc, resolution, depth) / decode(bits, resolution)                octree_np.py:10-45 / 47-112
(pn_kit.binary_array_to_byte_array, pn_kit.py:463-467, is the `bytes` by-product of ops.octree_encode.)

The reference calls these with numpy arrays it has just pulled off the device; the drop-ins take numpy arrays OR CUDA
tensors, run the coder on the GPU (one CTA per cloud) and return what the reference returns (numpy uint8 bit arrays /
float32 centres).  `encode_sampled` is the device-resident form used by pcc_b200.codec (no host round trip at all).
Only scale == 1 is supported -- the only value the reference ever passes.
"""
import numpy as np
import torch

from . import ops


def _to_cuda(a):
    if isinstance(a, torch.Tensor):
        if not a.is_cuda:
            a = a.cuda()
        return a.float()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def _check_scale(scale):
    if float(scale) != 1.0:
        raise NotImplementedError("pcc_b200 octree coder: only scale / resolution == 1 is supported (every reference call site)")


def _check_depths(depth):
    if bool((depth < 0).any()):
        raise ValueError("pcc_b200 octree coder: coordinates must lie in [0, 1) (octree_np.py:5-7)")


def encode_sampled(sampled_xyz, N, min_bpp, **by_products):
    """Device-resident encode_sampled_np: CUDA [B,S,3] in, dict of CUDA tensors out (see ops.octree_encode)."""
    return ops.octree_encode(sampled_xyz, N, min_bpp, 0, **by_products)


def encode_sampled_np(sampled_xyz, scale, N, min_bpp):
    """pn_kit.encode_sampled_np -> (codes: list of np.uint8 bit arrays, codebits: int)."""
    _check_scale(scale)
    r = ops.octree_encode(_to_cuda(sampled_xyz), N, min_bpp, 0)
    bits, nbits, depth = r["bits"].cpu().numpy(), r["nbits"].cpu().numpy(), r["depth"].cpu()
    _check_depths(depth)
    codes = [bits[b, :nbits[b]].copy() for b in range(bits.shape[0])]
    return codes, int(nbits.sum())


def _pack_codes(codes):
    n = np.array([len(c) for c in codes], dtype=np.int32)
    bits = np.zeros((len(codes), max(int(n.max()), 1)), dtype=np.uint8)
    for b, c in enumerate(codes):
        bits[b, :n[b]] = np.asarray(c, dtype=np.uint8)
    return torch.from_numpy(bits).cuda(), torch.from_numpy(n).cuda()


def decode_sampled_np(codes, scale):
    """pn_kit.decode_sampled_np -> np.float32 [B,64,3]: the reference decoder as written (octree_np.py:47-112)."""
    _check_scale(scale)
    bits, n = _pack_codes(codes)
    xyz, _, _ = ops.octree_decode(bits, n, mode=0, cap=64)
    return xyz.cpu().numpy()


def encode(pc, resolution, depth):
    """octree_np.encode(pc [S,3], resolution, depth) -> np.uint8 bit array."""
    _check_scale(resolution)
    if depth < 1:
        raise NotImplementedError("pcc_b200 octree coder: depth must be >= 1")
    r = ops.octree_encode(_to_cuda(pc)[None], 1, 0.0, int(depth))
    _check_depths(r["depth"].cpu())
    return r["bits"][0, :int(r["nbits"][0])].cpu().numpy()


def decode(bits, resolution):
    """octree_np.decode(bits, resolution) -> np.float32 [64,3] (as written in the reference)."""
    return decode_sampled_np([bits], resolution)[0]


def decode_inverse(codes, cap):
    """The decoder the reference lacks: leaf centres of every stream in stream order, np.float32 [B,cap,3] + counts."""
    bits, n = _pack_codes(codes)
    xyz, count, _ = ops.octree_decode(bits, n, mode=1, cap=cap)
    return xyz.cpu().numpy(), count.cpu().numpy()

