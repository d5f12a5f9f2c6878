"""ctypes front end of the CPU oracle (oracle/pcc_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
Every function takes and returns CPU numpy arrays (float32 / int64) and mirrors the signature of the reference
function it restates; see pcc_oracle.c for the reference file:line of each.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpcc_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i64p = ctypes.POINTER(ctypes.c_int64)
_f64p = ctypes.POINTER(ctypes.c_double)
_i64 = ctypes.c_int64
_u8p = ctypes.POINTER(ctypes.c_uint8)


def build(force=False):
    """Compile libpcc_oracle.so with the recipe in oracle/Makefile."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
            os.path.join(_HERE, "pcc_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={k: v for k, v in os.environ.items() if k != "CC"})
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.orc_num_threads.restype = ctypes.c_int
        L.orc_fps.argtypes = [_f32p, _i64, _i64, _i64, _i64p, ctypes.c_float, ctypes.c_int, _i64p, ctypes.c_int]
        L.orc_gather.argtypes = [_f32p, _i64p, _i64, _i64, _i64, _i64, _f32p]
        L.orc_knn.argtypes = [_f32p, _f32p, _i64, _i64, _i64, _i64, _f32p, _i64p, ctypes.c_int]
        L.orc_ball_query.argtypes = [_f32p, _f32p, _i64, _i64, _i64, _i64, ctypes.c_float, _i64p, _f32p,
                                     ctypes.c_int]
        L.orc_nn1.argtypes = [_f32p, _f32p, _i64, _i64, _i64, _f32p, _i64p, ctypes.c_int]
        L.orc_chamfer.argtypes = [_f32p, _f32p, _i64, _i64, _i64, _f32p, _i64p, _f32p, _i64p, _f64p, ctypes.c_int]
        L.orc_chamfer.restype = ctypes.c_double
        L.orc_chamfer_bwd.argtypes = [_f32p, _f32p, _i64p, _i64p, _i64, _i64, _i64, ctypes.c_float, _f32p, _f32p]
        L.orc_d1_psnr.argtypes = [_f32p, _i64, _f32p, _i64, _f64p]
        L.orc_d1_psnr.restype = ctypes.c_double
        L.orc_octree_quantise.argtypes = [_f32p, _i64, ctypes.c_double, ctypes.c_int, _f32p, _f32p]
        L.orc_octree_quantise.restype = _i64
        L.orc_octree_encode.argtypes = [_f32p, _i64, ctypes.c_double, ctypes.c_int, _u8p, _i64, _i64p]
        L.orc_octree_encode.restype = _i64
        L.orc_octree_encode_sampled.argtypes = [_f32p, _i64, ctypes.c_double, _i64, ctypes.c_double, _u8p, _i64,
                                                ctypes.POINTER(ctypes.c_int)]
        L.orc_octree_encode_sampled.restype = _i64
        L.orc_octree_decode_ref.argtypes = [_u8p, _i64, ctypes.c_double, _f32p]
        L.orc_bits_to_bytes.argtypes = [_u8p, _i64, _u8p]
        L.orc_bits_to_bytes.restype = _i64
        _u16p, _i16p = ctypes.POINTER(ctypes.c_uint16), ctypes.POINTER(ctypes.c_int16)
        L.orc_pmf_to_cdf_u16.argtypes = [_f32p, _i64, ctypes.c_int, _u16p]
        L.orc_range_encode.argtypes = [_u16p, _i16p, _i64, ctypes.c_int, _u8p, _i64]
        L.orc_range_encode.restype = _i64
        L.orc_range_decode.argtypes = [_u16p, _i64, ctypes.c_int, _u8p, _i64, _i16p]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _pf(a):
    return a.ctypes.data_as(_f32p)


def _pi(a):
    return a.ctypes.data_as(_i64p)


def num_threads():
    return int(lib().orc_num_threads())


def fps(xyz, npoint, start_idx=None, init_dist=1e10, pad_beyond_n=False, threads=1):
    """pn_kit.farthest_point_sample_batch (start_idx = the reference's randint draw, init 1e10) or PyTorch3D
    sample_farthest_points (start_idx=None, init FLT_MAX, pad_beyond_n=True)."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    out = np.empty((B, npoint), dtype=np.int64)
    si = None
    if start_idx is not None:
        si = np.ascontiguousarray(start_idx, dtype=np.int64)
    lib().orc_fps(_pf(xyz), B, N, npoint, _pi(si) if si is not None else None, np.float32(init_dist),
                  int(bool(pad_beyond_n)), _pi(out), threads)
    return out


FLT_MAX = float(np.finfo(np.float32).max)


def sample_farthest_points(points, K):
    """PyTorch3D semantics (pointnet_sa_module.py:10-13): returns (gathered points, idx); idx padded with -1."""
    idx = fps(points, K, None, FLT_MAX, True)
    pts = gather(points, np.maximum(idx, 0))
    pts[idx < 0] = 0.0  # masked_gather zeroes the padded rows
    return pts, idx


def gather(feat, idx):
    """index_points / knn_gather: feat [B,N,C], idx [B,...] -> [B,...,C]."""
    feat = _f32(feat)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    B, N, C = feat.shape
    M = int(np.prod(idx.shape[1:])) if idx.ndim > 1 else 1
    out = np.empty((B, M, C), dtype=np.float32)
    lib().orc_gather(_pf(feat), _pi(idx), B, N, C, M, _pf(out))
    return out.reshape(tuple(idx.shape) + (C,))


def knn_points(p1, p2, K, return_nn=False, threads=1):
    p1, p2 = _f32(p1), _f32(p2)
    B, P1, _ = p1.shape
    P2 = p2.shape[1]
    d = np.empty((B, P1, K), dtype=np.float32)
    i = np.empty((B, P1, K), dtype=np.int64)
    lib().orc_knn(_pf(p1), _pf(p2), B, P1, P2, K, _pf(d), _pi(i), threads)
    nn = gather(p2, i) if return_nn else None
    return d, i, nn


def ball_query(p1, p2, K, radius, threads=1):
    p1, p2 = _f32(p1), _f32(p2)
    B, P1, _ = p1.shape
    P2 = p2.shape[1]
    i = np.empty((B, P1, K), dtype=np.int64)
    d = np.empty((B, P1, K), dtype=np.float32)
    lib().orc_ball_query(_pf(p1), _pf(p2), B, P1, P2, K, np.float32(radius), _pi(i), _pf(d), threads)
    return d, i


def nn1(p1, p2, threads=1):
    p1, p2 = _f32(p1), _f32(p2)
    B, P1, _ = p1.shape
    P2 = p2.shape[1]
    d = np.empty((B, P1), dtype=np.float32)
    i = np.empty((B, P1), dtype=np.int64)
    lib().orc_nn1(_pf(p1), _pf(p2), B, P1, P2, _pf(d), _pi(i), threads)
    return d, i


def chamfer(x, y, threads=1):
    """Returns (loss, per_cloud[B], dx[B,P1], ix, dy[B,P2], iy)."""
    x, y = _f32(x), _f32(y)
    B, P1, _ = x.shape
    P2 = y.shape[1]
    dx = np.empty((B, P1), dtype=np.float32)
    ix = np.empty((B, P1), dtype=np.int64)
    dy = np.empty((B, P2), dtype=np.float32)
    iy = np.empty((B, P2), dtype=np.int64)
    pc = np.empty((B,), dtype=np.float64)
    loss = lib().orc_chamfer(_pf(x), _pf(y), B, P1, P2, _pf(dx), _pi(ix), _pf(dy), _pi(iy),
                             pc.ctypes.data_as(_f64p), threads)
    return float(loss), pc, dx, ix, dy, iy


def chamfer_bwd(x, y, ix, iy, grad=1.0):
    x, y = _f32(x), _f32(y)
    B, P1, _ = x.shape
    P2 = y.shape[1]
    gx = np.zeros_like(x)
    gy = np.zeros_like(y)
    ix = np.ascontiguousarray(ix, dtype=np.int64)
    iy = np.ascontiguousarray(iy, dtype=np.int64)
    lib().orc_chamfer_bwd(_pf(x), _pf(y), _pi(ix), _pi(iy), B, P1, P2, np.float32(grad), _pf(gx), _pf(gy))
    return gx, gy


def d1_psnr(orig, recon):
    """eval.py:43-98 point-to-point PSNR (float64), returns (psnr_db, mse)."""
    orig, recon = _f32(orig), _f32(recon)
    mse = ctypes.c_double(0.0)
    psnr = lib().orc_d1_psnr(_pf(orig), orig.shape[0], _pf(recon), recon.shape[0], ctypes.byref(mse))
    return float(psnr), float(mse.value)


# ---- octree centre coding (octree_np.py, pn_kit.py:380-475) -------------------------------------------------------
def _pu8(a):
    return a.ctypes.data_as(_u8p)


def octree_quantise(pc, resolution, depth):
    """octree_np.getDecodeFromPc for one [S,3] cloud: returns (snapped [S,3] in input order, unique rows [U,3])."""
    pc = _f32(pc)
    S = pc.shape[0]
    snapped = np.empty((S, 3), np.float32)
    uniq = np.empty((max(S, 1), 3), np.float32)
    n = lib().orc_octree_quantise(_pf(pc), S, float(resolution), int(depth), _pf(snapped), _pf(uniq))
    return snapped, uniq[:n].copy()


def octree_encode(pc, resolution, depth):
    """octree_np.encode(pc, resolution, depth) -> uint8 bit array."""
    pc = _f32(pc)
    S = pc.shape[0]
    cap = 1 + 8 * (depth + 1) * max(S, 1) + 64
    bits = np.zeros(cap, np.uint8)
    n = lib().orc_octree_encode(_pf(pc), S, float(resolution), int(depth), _pu8(bits), cap, None)
    assert n >= 0
    return bits[:n].copy()


def encode_sampled_np(sampled_xyz, scale, N, min_bpp):
    """pn_kit.encode_sampled_np -> (codes list, total bits, depths)."""
    sampled_xyz = _f32(sampled_xyz)
    codes, depths, total = [], [], 0
    S = sampled_xyz.shape[1]
    cap = 1 + 8 * 17 * max(S, 1) + 64
    for pc in sampled_xyz:
        pc = np.ascontiguousarray(pc)
        bits = np.zeros(cap, np.uint8)
        d = ctypes.c_int(0)
        n = lib().orc_octree_encode_sampled(_pf(pc), S, float(scale), int(N), float(min_bpp), _pu8(bits), cap,
                                            ctypes.byref(d))
        assert n >= 0
        codes.append(bits[:n].copy())
        depths.append(d.value)
        total += int(n)
    return codes, total, depths


def octree_decode_ref(bits, resolution=1.0):
    """octree_np.decode exactly as the reference wrote it (first 8 bits -> depth-1 octant centres, padded to 64)."""
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    out = np.empty((64, 3), np.float32)
    lib().orc_octree_decode_ref(_pu8(bits), bits.shape[0], float(resolution), _pf(out))
    return out


def bits_to_bytes(bits):
    """pn_kit.binary_array_to_byte_array."""
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    out = np.zeros((bits.shape[0] + 7) // 8, np.uint8)
    n = lib().orc_bits_to_bytes(_pu8(bits), bits.shape[0], _pu8(out))
    return out[:n]


def octree_stream_centres(pc, depth, rows):
    """The centres a correct decoder recovers from octree_encode(pc, 1, depth): the distinct snapped cells in the
    stream's own order (the reference's DFS pops children 7..0 => descending (x, y, z)-interleaved cell code), the last
    one repeated up to `rows` rows (the reference's padding rule, octree_np.py:101-105).  numpy, small inputs."""
    _, u = octree_quantise(pc, 1.0, depth)
    cells = np.floor(u * np.float32(2.0 ** depth)).astype(np.int64)
    code = np.zeros(len(u), np.int64)
    for lvl in range(depth - 1, -1, -1):
        code = (code << 3) | (((cells[:, 0] >> lvl) & 1) << 2) | (((cells[:, 1] >> lvl) & 1) << 1) | ((cells[:, 2] >> lvl) & 1)
    u = u[np.argsort(-code, kind="stable")]
    if len(u) < rows:
        u = np.concatenate([u, np.repeat(u[-1:], rows - len(u), axis=0)])
    return u[:rows]


# ---- the remaining eval.py metrics ---------------------------------------------------------------------
def calc_uc(input_pc, decomp_pc, region=1024):
    """eval.py:127-151 calc_uc: 1024-NN region of point 0 (knn_points), distance of every region point to its nearest
    other region point (the reference takes column 1 of a topk over torch.cdist), np.var ratio."""
    def region_dist(pc):
        pc = _f32(pc)
        _, _, nn = knn_points(pc[None, :1], pc[None], region, True)
        reg = nn[0, 0] - pc[0]                                   # eval.py:134 recentring
        d, _, _ = knn_points(reg[None], reg[None], 2, False)
        return np.sqrt(d[0, :, 1].astype(np.float64))
    return float(np.var(region_dist(decomp_pc)) / np.var(region_dist(input_pc)))


def estimate_normals(points, knn=30):
    """Open3D estimate_normals(KDTreeSearchParamKNN(knn)), eval.py:58-59 (Open3D is not installed here: PARITY UNPINNED;
    restated from its published algorithm -- covariance of the knn nearest points, eigenvector of the smallest eigenvalue)."""
    points = _f32(points)
    _, idx, _ = knn_points(points[None], points[None], knn, False)
    nb = points[idx[0]].astype(np.float64)                        # [N, knn, 3]
    mean = nb.mean(axis=1, keepdims=True)
    cov = np.einsum("nki,nkj->nij", nb, nb) / knn - np.einsum("nki,nkj->nij", mean, mean)
    _, v = np.linalg.eigh(cov)
    return v[:, :, 0]


def p2plane_psnr(orig, recon, knn=30):
    """eval.py:43-98 compute_p2point_p2plane_psnr, p2plane half (float64): returns (psnr_db, mse)."""
    orig, recon = _f32(orig), _f32(recon)
    normals = estimate_normals(orig, knn)
    _, ix = nn1(recon[None], orig[None])
    diff = recon.astype(np.float64) - orig[ix[0]].astype(np.float64)
    e = np.einsum("ni,ni->n", diff, normals[ix[0]]) ** 2
    mse = float(e.mean())
    diag = np.linalg.norm(orig.max(axis=0).astype(np.float64) - orig.min(axis=0).astype(np.float64))
    return (float(10 * np.log10(diag ** 2 / mse)) if mse > 0 else float("inf")), mse


# ---- entropy stage (pn_kit.pmf_to_cdf + torchac's coder; torchac is absent: PARITY UNPINNED) ----------------------------
def pmf_to_cdf_u16(pmf):
    """pn_kit.pmf_to_cdf followed by torchac's _convert_to_int_and_normalize(needs_normalization=True): [..., L] -> uint16 [..., L+1]."""
    pmf = _f32(pmf)
    L = pmf.shape[-1]
    out = np.empty(pmf.shape[:-1] + (L + 1,), np.uint16)
    lib().orc_pmf_to_cdf_u16(_pf(pmf), pmf.size // L, L, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)))
    return out


def range_encode(cdf_u16, sym):
    """torchac.encode_float_cdf's coder on an integer CDF: cdf uint16 [n, Lp], sym int16 [n] -> bytes."""
    cdf = np.ascontiguousarray(cdf_u16, np.uint16).reshape(-1, cdf_u16.shape[-1])
    sym = np.ascontiguousarray(sym, np.int16).reshape(-1)
    cap = 4 * sym.size + 64
    out = np.zeros(cap, np.uint8)
    n = lib().orc_range_encode(cdf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), sym.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)),
                               sym.size, cdf.shape[1], _pu8(out), cap)
    assert n <= cap
    return out[:n].tobytes()


def range_decode(cdf_u16, data):
    cdf = np.ascontiguousarray(cdf_u16, np.uint16).reshape(-1, cdf_u16.shape[-1])
    buf = np.frombuffer(data, np.uint8).copy()
    sym = np.empty(cdf.shape[0], np.int16)
    lib().orc_range_decode(cdf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), cdf.shape[0], cdf.shape[1], _pu8(buf), buf.size,
                           sym.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)))
    return sym
