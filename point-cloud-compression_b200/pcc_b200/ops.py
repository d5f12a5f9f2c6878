"""Tensor-level host functions over the C ABI: the B200 implementation of the reference's geometric ops.

Every function takes CUDA float32 tensors, enqueues hand-written sm_100a kernels on the current torch stream and
returns torch tensors; nothing here computes on the CPU.  Signatures follow the reference call sites cited in
include/pcc_b200.h.
"""
import collections

import os

import torch

from . import _lib

_KNN = collections.namedtuple("KNN", "dists idx knn")  # same fields as pytorch3d.ops.knn._KNN
FLT_MAX = 3.4028234663852886e38


def _cuda_f32(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"pcc_b200: {name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _check_pair(p1, p2):
    if p1.dim() != 3 or p2.dim() != 3:
        raise ValueError("pcc_b200: point tensors must have shape [B, P, 3]")
    if p1.shape[0] != p2.shape[0]:
        raise ValueError("pts1 and pts2 must have the same batch dimension.")
    if p1.shape[2] != 3 or p2.shape[2] != 3:
        raise ValueError("pcc_b200: only 3-D points are supported (pts1 and pts2 must have the same point dimension 3).")


def fps(xyz, npoint, start_idx=None, init_dist=1e10, return_xyz=False, quant_cube=0.0):
    """Farthest point sampling.  start_idx [B] int64 (device) or None (= start at 0).  Returns int64 [B,npoint];
    with return_xyz also the sampled points [B,npoint,3] (snapped to the octree grid when quant_cube > 0)."""
    lib = _lib.load()
    xyz = _cuda_f32(xyz, "xyz")
    B, N, C = xyz.shape
    if C != 3:
        raise ValueError("pcc_b200.fps: xyz must be [B, N, 3]")
    out = torch.empty((B, npoint), dtype=torch.int64, device=xyz.device)
    out_xyz = torch.empty((B, npoint, 3), dtype=torch.float32, device=xyz.device) if return_xyz else None
    if start_idx is not None:
        start_idx = start_idx.to(device=xyz.device, dtype=torch.int64).contiguous()
    with torch.cuda.device(xyz.device):
        ws_bytes = lib.pcc_fps_workspace_bytes(B, N, npoint)
        ws = torch.empty((max(ws_bytes, 1),), dtype=torch.uint8, device=xyz.device) if ws_bytes else None
        _lib.check(lib.pcc_fps_f32(_ptr(xyz), B, N, npoint, _ptr(start_idx), init_dist, _ptr(out), _ptr(out_xyz),
                                   float(quant_cube), _ptr(ws), _stream()), "pcc_fps_f32")
    return (out, out_xyz) if return_xyz else out


def normalize(xyz, margin=0.01):
    """pn_kit.normalize per cloud (pn_kit.py:47-60): (normalised [B,N,3], center [B,3], longest [B], bbox [B,6])."""
    lib = _lib.load()
    xyz = _cuda_f32(xyz, "xyz")
    B, N, _ = xyz.shape
    out = torch.empty_like(xyz)
    center = torch.empty((B, 3), dtype=torch.float32, device=xyz.device)
    longest = torch.empty((B,), dtype=torch.float32, device=xyz.device)
    bbox = torch.empty((B, 6), dtype=torch.float32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        _lib.check(lib.pcc_normalize_f32(_ptr(xyz), B, N, float(margin), _ptr(out), _ptr(center), _ptr(longest), _ptr(bbox),
                                         _stream()), "pcc_normalize_f32")
    return out, center, longest, bbox


def assemble(patches, centres, patch_scale, center=None, longest=None, margin=0.01):
    """decompress.py:104-116: patches [B*S,k,3] / patch_scale + centres [B,S,3] (+ denormalise) -> [B, S*k, 3]."""
    lib = _lib.load()
    patches, centres = _cuda_f32(patches, "patches"), _cuda_f32(centres, "centres")
    B, S, _ = centres.shape
    k = patches.shape[1]
    out = torch.empty((B, S * k, 3), dtype=torch.float32, device=patches.device)
    if center is not None:
        center, longest = _cuda_f32(center, "center"), _cuda_f32(longest, "longest")
    with torch.cuda.device(patches.device):
        _lib.check(lib.pcc_assemble_f32(_ptr(patches), _ptr(centres), _ptr(center), _ptr(longest), B, S, k,
                                        float(patch_scale), float(margin), _ptr(out), _stream()), "pcc_assemble_f32")
    return out


_KNN_GRID_MIN_POINTS = 65536      # candidate clouds at least this large go through the grid form of the search
_KNN_NO_GRID = bool(os.environ.get("PCC_KNN_NO_GRID"))   # A/B switch for the measurements in profiles/


def knn(p1, p2, K, return_nn=False, centre_sub=False, nn_scale=1.0, nn_only=False, grid=None):
    """K nearest neighbours: (dists [B,P1,K] squared, idx int64 [B,P1,K], nn [B,P1,K,3] or None).
    nn_only=True skips the distance / index outputs (returns None for them): the grouping calls of the network bodies
    use only the gathered, recentred neighbours.  grid: None = the grid form of the search for scene-scale p2 (>= 65536 points),
    True / False force / forbid it (tests: both forms give identical results)."""
    lib = _lib.load()
    p1, p2 = _cuda_f32(p1, "p1"), _cuda_f32(p2, "p2")
    _check_pair(p1, p2)
    B, P1, _ = p1.shape
    P2 = p2.shape[1]
    if nn_only and not return_nn:
        raise ValueError("pcc_b200.knn: nn_only requires return_nn")
    d = None if nn_only else torch.empty((B, P1, K), dtype=torch.float32, device=p1.device)
    i = None if nn_only else torch.empty((B, P1, K), dtype=torch.int64, device=p1.device)
    nn = torch.empty((B, P1, K, 3), dtype=torch.float32, device=p1.device) if return_nn else None
    if B == 0 or P1 == 0:   # nothing to search for: empty results, no launch (empty tensors have no storage to point at)
        return d, i, nn
    with torch.cuda.device(p1.device):
        if (P2 >= _KNN_GRID_MIN_POINTS and not _KNN_NO_GRID) if grid is None else grid:
            # scene scale: the same search through a uniform grid over p2 (identical results; workspace owned by this call)
            ws = torch.empty((lib.pcc_knn_grid_workspace_bytes(B, P2),), dtype=torch.uint8, device=p1.device)
            _lib.check(lib.pcc_knn_grid_f32(_ptr(p1), _ptr(p2), B, P1, P2, K, _ptr(d), _ptr(i), _ptr(nn), int(centre_sub),
                                            float(nn_scale), _ptr(ws), _stream()), "pcc_knn_grid_f32")
        else:
            _lib.check(lib.pcc_knn_f32(_ptr(p1), _ptr(p2), B, P1, P2, K, _ptr(d), _ptr(i), _ptr(nn), int(centre_sub),
                                       float(nn_scale), _stream()), "pcc_knn_f32")
    return d, i, nn


def knn_patch_u8(patches, K):
    """In-patch kNN table as bytes: patches [BS, P, 3] (K <= P <= 256, K in {8, 16}) -> uint8 [BS, P, K], the K nearest points of
    every point inside its own patch in (d2, idx) order (pn_kit.py:190 with S == N); feeds mlp_ops.sa_chain_indexed."""
    lib = _lib.load()
    patches = _cuda_f32(patches, "patches")
    BS, P, _ = patches.shape
    out = torch.empty((BS, P, K), dtype=torch.uint8, device=patches.device)
    if BS == 0:
        return out
    with torch.cuda.device(patches.device):
        _lib.check(lib.pcc_knn_patch_u8(_ptr(patches), BS, P, K, _ptr(out), _stream()), "pcc_knn_patch_u8")
    return out


def ball_query(p1, p2, K, radius, return_dists=True):
    lib = _lib.load()
    p1, p2 = _cuda_f32(p1, "p1"), _cuda_f32(p2, "p2")
    _check_pair(p1, p2)
    B, P1, _ = p1.shape
    P2 = p2.shape[1]
    i = torch.empty((B, P1, K), dtype=torch.int64, device=p1.device)
    d = torch.empty((B, P1, K), dtype=torch.float32, device=p1.device) if return_dists else None
    if B == 0 or P1 == 0:
        return d, i
    with torch.cuda.device(p1.device):
        _lib.check(lib.pcc_ball_query_f32(_ptr(p1), _ptr(p2), B, P1, P2, K, float(radius), _ptr(i), _ptr(d), _stream()),
                   "pcc_ball_query_f32")
    return d, i


class _Gather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, idx):
        lib = _lib.load()
        B, N, C = feat.shape
        M = idx.numel() // max(B, 1)
        out = torch.empty(tuple(idx.shape) + (C,), dtype=torch.float32, device=feat.device)
        if out.numel() > 0:
            with torch.cuda.device(feat.device):
                _lib.check(lib.pcc_gather_f32(_ptr(feat), _ptr(idx), B, N, C, M, _ptr(out), _stream()), "pcc_gather_f32")
        ctx.save_for_backward(idx)
        ctx.shape = (B, N, C, M)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        (idx,) = ctx.saved_tensors
        B, N, C, M = ctx.shape
        grad_out = grad_out.contiguous().float()
        grad_feat = torch.zeros((B, N, C), dtype=torch.float32, device=grad_out.device)
        if grad_out.numel() > 0:
            with torch.cuda.device(grad_out.device):
                _lib.check(lib.pcc_gather_bwd_f32(_ptr(grad_out), _ptr(idx), B, N, C, M, _ptr(grad_feat), _stream()),
                           "pcc_gather_bwd_f32")
        return grad_feat, None


def gather(feat, idx):
    """out[b, ..., :] = feat[b, idx[b, ...], :]; differentiable w.r.t. feat.  idx must lie in [0, N)."""
    feat = _cuda_f32(feat, "feat")
    if feat.dim() != 3 or idx.shape[0] != feat.shape[0]:
        raise ValueError("pcc_b200.gather: feat must be [B,N,C] and idx [B,...]")
    idx = idx.to(device=feat.device, dtype=torch.int64).contiguous()
    return _Gather.apply(feat, idx)


def nn1(p1, p2, return_idx=True):
    """Nearest neighbour of each p1 point in p2: (d2 [B,P1], idx int64 [B,P1] or None)."""
    lib = _lib.load()
    p1, p2 = _cuda_f32(p1, "p1"), _cuda_f32(p2, "p2")
    _check_pair(p1, p2)
    B, P1, _ = p1.shape
    P2 = p2.shape[1]
    d = torch.empty((B, P1), dtype=torch.float32, device=p1.device)
    i = torch.empty((B, P1), dtype=torch.int64, device=p1.device) if return_idx else None
    if B == 0 or P1 == 0:
        return d, i
    with torch.cuda.device(p1.device):
        ws = torch.empty((max(lib.pcc_nn1_workspace_bytes(B, P1, P2), 8),), dtype=torch.uint8, device=p1.device)
        _lib.check(lib.pcc_nn1_f32(_ptr(p1), _ptr(p2), B, P1, P2, _ptr(d), _ptr(i), _ptr(ws), _stream()), "pcc_nn1_f32")
    return d, i


def chamfer_forward(x, y, want_idx=True):
    """Returns dict(loss [scalar tensor], per_cloud [B], dx, ix, dy, iy)."""
    lib = _lib.load()
    x, y = _cuda_f32(x, "x"), _cuda_f32(y, "y")
    _check_pair(x, y)
    B, P1, _ = x.shape
    P2 = y.shape[1]
    dev = x.device
    dx = torch.empty((B, P1), dtype=torch.float32, device=dev)
    dy = torch.empty((B, P2), dtype=torch.float32, device=dev)
    ix = torch.empty((B, P1), dtype=torch.int64, device=dev) if want_idx else None
    iy = torch.empty((B, P2), dtype=torch.int64, device=dev) if want_idx else None
    per_cloud = torch.empty((B,), dtype=torch.float32, device=dev)
    loss = torch.empty((1,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ws = torch.empty((max(lib.pcc_chamfer_workspace_bytes(B, P1, P2), 8),), dtype=torch.uint8, device=dev)
        _lib.check(lib.pcc_chamfer_fwd_f32(_ptr(x), _ptr(y), B, P1, P2, _ptr(dx), _ptr(ix), _ptr(dy), _ptr(iy),
                                           _ptr(per_cloud), _ptr(loss), _ptr(ws), _stream()), "pcc_chamfer_fwd_f32")
    return dict(loss=loss.reshape(()), per_cloud=per_cloud, dx=dx, ix=ix, dy=dy, iy=iy)


def eval_metrics(dx, per_cloud, bbox):
    """[B,3] float64 (normalised Chamfer, D1 PSNR dB, D1 MSE) from chamfer_forward's dx / per_cloud and the original's
    bbox [B,6] (eval.py:84,88-92,199-205)."""
    lib = _lib.load()
    B, P1 = dx.shape
    bbox = bbox.detach().float().contiguous()
    out = torch.empty((B, 3), dtype=torch.float64, device=dx.device)
    with torch.cuda.device(dx.device):
        _lib.check(lib.pcc_eval_metrics_f32(_ptr(dx), _ptr(per_cloud), _ptr(bbox), B, P1, _ptr(out), _stream()),
                   "pcc_eval_metrics_f32")
    return out


class _Chamfer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        r = chamfer_forward(x, y, want_idx=True)
        ctx.save_for_backward(x, y, r["ix"], r["iy"])
        return r["loss"]

    @staticmethod
    def backward(ctx, grad):
        lib = _lib.load()
        x, y, ix, iy = ctx.saved_tensors
        x, y = x.contiguous().float(), y.contiguous().float()
        B, P1, _ = x.shape
        P2 = y.shape[1]
        gx = torch.empty_like(x)
        gy = torch.empty_like(y)
        g = grad.detach().to(device=x.device, dtype=torch.float32).reshape(1).contiguous()
        with torch.cuda.device(x.device):
            _lib.check(lib.pcc_chamfer_bwd_f32(_ptr(x), _ptr(y), _ptr(ix), _ptr(iy), B, P1, P2, _ptr(g), _ptr(gx),
                                               _ptr(gy), _stream()), "pcc_chamfer_bwd_f32")
        return gx, gy


def chamfer(x, y):
    """Differentiable Chamfer distance (mean/mean), scalar tensor."""
    return _Chamfer.apply(_cuda_f32(x, "x"), _cuda_f32(y, "y"))


def octree_encode(centres, n_points, min_bpp=0.0, fixed_depth=0, want_bytes=False, want_quant=False, want_rec_ref=False,
                  want_stream_xyz=False):
    """Octree coding of patch centres [B,S,3] in [0,1) (octree_np.encode + pn_kit.encode_sampled_np's depth search).
    Returns dict(bits uint8 [B,max_bits], nbits int32 [B], depth int32 [B], and the requested by-products: bytes,
    quant [B,S,3], rec_ref [B,64,3], stream_xyz [B,S,3]).  Everything stays on the device; nothing is synchronised."""
    lib = _lib.load()
    centres = _cuda_f32(centres, "centres")
    if centres.dim() != 3 or centres.shape[2] != 3:
        raise ValueError("pcc_b200.octree_encode: centres must be [B, S, 3]")
    B, S, _ = centres.shape
    dev = centres.device
    max_bits = 1 + 8 * fixed_depth * S if fixed_depth > 0 else lib.pcc_octree_max_bits(S)
    bits = torch.empty((B, max_bits), dtype=torch.uint8, device=dev)
    nbits = torch.empty((B,), dtype=torch.int32, device=dev)
    depth = torch.empty((B,), dtype=torch.int32, device=dev)
    by = torch.empty((B, (max_bits + 7) // 8), dtype=torch.uint8, device=dev) if want_bytes else None
    quant = torch.empty((B, S, 3), dtype=torch.float32, device=dev) if want_quant else None
    rec = torch.empty((B, 64, 3), dtype=torch.float32, device=dev) if want_rec_ref else None
    sx = torch.empty((B, S, 3), dtype=torch.float32, device=dev) if want_stream_xyz else None
    with torch.cuda.device(dev):
        _lib.check(lib.pcc_octree_encode_f32(_ptr(centres), B, S, int(n_points), float(min_bpp), int(fixed_depth), _ptr(bits),
                                             max_bits, _ptr(nbits), _ptr(depth), _ptr(by), _ptr(quant), _ptr(rec), _ptr(sx),
                                             _stream()), "pcc_octree_encode_f32")
    return dict(bits=bits, nbits=nbits, depth=depth, bytes=by, quant=quant, rec_ref=rec, stream_xyz=sx)


def octree_decode(bits, nbits, mode=1, cap=64):
    """bits uint8 [B,max_bits] (one byte per bit), nbits int32 [B].  mode 0: the reference's octree_np.decode as written
    (first 8 bits -> depth-1 octant centres padded to 64 rows); mode 1: the inverse of octree_encode.
    Returns (xyz [B,cap,3], count int32 [B], depth int32 [B])."""
    lib = _lib.load()
    if not bits.is_cuda or bits.dtype != torch.uint8 or bits.dim() != 2:
        raise RuntimeError("pcc_b200.octree_decode: bits must be a CUDA uint8 [B, max_bits] tensor (there is no CPU path)")
    bits = bits.contiguous()
    nbits = nbits.to(device=bits.device, dtype=torch.int32).contiguous()
    B, max_bits = bits.shape
    out = torch.empty((B, cap, 3), dtype=torch.float32, device=bits.device)
    count = torch.empty((B,), dtype=torch.int32, device=bits.device)
    depth = torch.empty((B,), dtype=torch.int32, device=bits.device)
    with torch.cuda.device(bits.device):
        _lib.check(lib.pcc_octree_decode_f32(_ptr(bits), _ptr(nbits), B, max_bits, int(mode), int(cap), _ptr(out), _ptr(count),
                                             _ptr(depth), _stream()), "pcc_octree_decode_f32")
    return out, count, depth


def estimate_normals(points, knn_k=30):
    """Open3D estimate_normals(KDTreeSearchParamKNN(knn)) as eval.py:58-59 uses it: [B,N,3] -> unit normals [B,N,3]
    (PCA of the knn nearest points, the point itself included; sign arbitrary)."""
    lib = _lib.load()
    points = _cuda_f32(points, "points")
    B, N, _ = points.shape
    _, _, nn = knn(points, points, knn_k, return_nn=True, nn_only=True)
    out = torch.empty((B, N, 3), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        _lib.check(lib.pcc_normals_pca_f32(_ptr(nn), B * N, knn_k, _ptr(out), _stream()), "pcc_normals_pca_f32")
    return out


def p2plane_psnr(recon, orig, normals=None, ix=None, bbox=None, knn_k=30):
    """eval.py:43-98 compute_p2point_p2plane_psnr, the p2plane half: returns float64 [B,2] = (mse, psnr dB).
    normals / ix / bbox are computed here when not supplied (normals of `orig`, 1-NN of recon in orig, bbox of orig)."""
    lib = _lib.load()
    recon, orig = _cuda_f32(recon, "recon"), _cuda_f32(orig, "orig")
    _check_pair(recon, orig)
    B, P1, _ = recon.shape
    P2 = orig.shape[1]
    if normals is None:
        normals = estimate_normals(orig, knn_k)
    if ix is None:
        _, ix = nn1(recon, orig)
    if bbox is None:
        bbox = torch.cat((orig.amin(dim=1), orig.amax(dim=1)), dim=1)
    bbox = bbox.float().contiguous()
    ix = ix.to(torch.int64).contiguous()
    out = torch.empty((B, 2), dtype=torch.float64, device=recon.device)
    with torch.cuda.device(recon.device):
        _lib.check(lib.pcc_p2plane_f32(_ptr(recon), _ptr(orig), _ptr(ix), _ptr(normals.contiguous()), _ptr(bbox), B, P1, P2,
                                       _ptr(out), _stream()), "pcc_p2plane_f32")
    return out


def uniformity_coefficient(input_pc, decomp_pc, region=1024):
    """eval.py:127-151 calc_uc for a batch: the 1024 nearest points of point 0 of each cloud, every region point's distance
    to its nearest other region point, var(decompressed) / var(input).  Returns float64 [B]."""
    lib = _lib.load()
    input_pc, decomp_pc = _cuda_f32(input_pc, "input_pc"), _cuda_f32(decomp_pc, "decomp_pc")
    B = input_pc.shape[0]
    if input_pc.shape[1] < region or decomp_pc.shape[1] < region:
        raise ValueError("pcc_b200.uniformity_coefficient: clouds need at least `region` points")
    d2 = []
    for pc in (input_pc, decomp_pc):
        _, _, reg = knn(pc[:, :1].contiguous(), pc, region, return_nn=True, centre_sub=True, nn_only=True)  # KNN_Region incl.
        reg = reg.view(B, region, 3)                                                        # the recentring of eval.py:134
        dd, _, _ = knn(reg, reg, 2)
        d2.append(dd[:, :, 1].contiguous())
    out = torch.empty((B,), dtype=torch.float64, device=input_pc.device)
    with torch.cuda.device(input_pc.device):
        _lib.check(lib.pcc_uc_f32(_ptr(d2[0]), _ptr(d2[1]), B, region, _ptr(out), _stream()), "pcc_uc_f32")
    return out


# ---- entropy stage (pn_kit.pmf_to_cdf + torchac's coder) ----------------------------------------------------------------
def pmf_to_cdf_u16(pmf):
    """pmf [..., L] (CUDA fp32) -> uint16 CDF [..., L + 1]: pn_kit.pmf_to_cdf + torchac's 16-bit normalisation."""
    lib = _lib.load()
    pmf = _cuda_f32(pmf, "pmf")
    L = pmf.shape[-1]
    out = torch.empty(pmf.shape[:-1] + (L + 1,), dtype=torch.uint16, device=pmf.device)
    if pmf.numel():
        with torch.cuda.device(pmf.device):
            _lib.check(lib.pcc_pmf_to_cdf_u16(_ptr(pmf), pmf.numel() // L, L, _ptr(out), _stream()), "pcc_pmf_to_cdf_u16")
    return out


def cdf_to_u16(cdf_float):
    """float CDF [..., Lp] (what pn_kit.pmf_to_cdf returns) -> uint16 CDF, torchac's needs_normalization=True conversion."""
    lib = _lib.load()
    cdf_float = _cuda_f32(cdf_float, "cdf_float")
    Lp = cdf_float.shape[-1]
    out = torch.empty(cdf_float.shape, dtype=torch.uint16, device=cdf_float.device)
    if cdf_float.numel():
        with torch.cuda.device(cdf_float.device):
            _lib.check(lib.pcc_cdf_to_u16(_ptr(cdf_float), cdf_float.numel() // Lp, Lp, _ptr(out), _stream()), "pcc_cdf_to_u16")
    return out


def range_encode(cdf_u16, sym):
    """cdf uint16 [B, n, Lp], sym int16 [B, n] (CUDA) -> (bytes uint8 [B, cap], nbytes int32 [B]); one stream per cloud.
    cap = 2 n + 8 bytes bounds any stream over valid 16-bit CDFs (>= 1 count per symbol => <= 16 bits per symbol); a stream
    that would exceed it (a zero-width CDF interval) reports nbytes > cap and is truncated: callers that bring nbytes to the
    host check it with `check_stream_sizes`."""
    lib = _lib.load()
    if not cdf_u16.is_cuda or cdf_u16.dtype != torch.uint16 or cdf_u16.dim() != 3:
        raise RuntimeError("pcc_b200.range_encode: cdf must be a CUDA uint16 [B, n, Lp] tensor (there is no CPU path)")
    cdf_u16 = cdf_u16.contiguous()
    B, n, Lp = cdf_u16.shape
    sym = sym.to(device=cdf_u16.device, dtype=torch.int16).reshape(B, n).contiguous()
    cap = 2 * n + 8
    out = torch.empty((B, cap), dtype=torch.uint8, device=cdf_u16.device)
    nbytes = torch.empty((B,), dtype=torch.int32, device=cdf_u16.device)
    with torch.cuda.device(cdf_u16.device):
        _lib.check(lib.pcc_range_encode_u16(_ptr(cdf_u16), _ptr(sym), B, n, Lp, _ptr(out), cap, _ptr(nbytes), _stream()),
                   "pcc_range_encode_u16")
    return out, nbytes


def check_stream_sizes(nbytes_host, cap):
    """Raise when a coded stream did not fit its buffer (nbytes counts the bytes the coder produced, stored or not)."""
    import numpy as np
    nb = np.asarray(nbytes_host)
    if nb.size and int(nb.max()) > int(cap):
        raise RuntimeError(f"pcc_b200.range_encode: a stream needs {int(nb.max())} bytes but the buffer holds {int(cap)} "
                           "(invalid CDF: a symbol with an empty interval?)")


def range_decode(cdf_u16, data, nbytes):
    """Inverse of range_encode: returns sym int16 [B, n]."""
    lib = _lib.load()
    if not cdf_u16.is_cuda or cdf_u16.dtype != torch.uint16 or cdf_u16.dim() != 3:
        raise RuntimeError("pcc_b200.range_decode: cdf must be a CUDA uint16 [B, n, Lp] tensor (there is no CPU path)")
    cdf_u16 = cdf_u16.contiguous()
    B, n, Lp = cdf_u16.shape
    data = data.to(device=cdf_u16.device, dtype=torch.uint8).reshape(B, -1).contiguous()
    nbytes = nbytes.to(device=cdf_u16.device, dtype=torch.int32).contiguous()
    sym = torch.empty((B, n), dtype=torch.int16, device=cdf_u16.device)
    with torch.cuda.device(cdf_u16.device):
        _lib.check(lib.pcc_range_decode_u16(_ptr(cdf_u16), _ptr(data), _ptr(nbytes), B, n, Lp, data.shape[1], _ptr(sym), _stream()),
                   "pcc_range_decode_u16")
    return sym


def quantise_latent(raw, spread, kpad=0):
    """AE.py:42-45: (latent, latent_q[, latent_q as bf16 rows zero padded to kpad columns]) from the encoder output [rows, d]."""
    lib = _lib.load()
    raw = _cuda_f32(raw, "raw")
    rows, d = raw.shape
    latent = torch.empty_like(raw)
    q = torch.empty_like(raw)
    qb = torch.empty((rows, kpad), dtype=torch.bfloat16, device=raw.device) if kpad else None
    if rows:
        with torch.cuda.device(raw.device):
            _lib.check(lib.pcc_quantise_latent_f32(_ptr(raw), rows, d, kpad if kpad else d, float(spread), _ptr(latent), _ptr(q),
                                                   _ptr(qb), _stream()), "pcc_quantise_latent_f32")
    return latent, q, qb
