"""ctypes binding of libpcc_b200.so (the C ABI declared in include/pcc_b200.h).

There is no CPU fallback: if the shared library is missing the import fails loudly, and every op raises when
handed a non-CUDA tensor.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PCC_B200_LIB: development aid -- load another build of the same library (A/B runs of a kernel variant)
LIB_PATH = os.environ.get("PCC_B200_LIB") or os.path.join(_HERE, "libpcc_b200.so")

_vp = ctypes.c_void_p
_i = ctypes.c_int
_i64 = ctypes.c_int64
_f = ctypes.c_float

# name -> (restype, argtypes); mirrors include/pcc_b200.h one to one
SIGNATURES = {
    "pcc_version": (_i, []),
    "pcc_last_error_string": (ctypes.c_char_p, []),
    "pcc_launch_count": (_i64, []),
    "pcc_fps_workspace_bytes": (_i64, [_i, _i, _i]),
    "pcc_fps_f32": (_i, [_vp, _i, _i, _i, _vp, _f, _vp, _vp, _f, _vp, _vp]),
    "pcc_knn_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _f, _vp]),
    "pcc_ball_query_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp]),
    "pcc_gather_f32": (_i, [_vp, _vp, _i, _i, _i, _i64, _vp, _vp]),
    "pcc_gather_bwd_f32": (_i, [_vp, _vp, _i, _i, _i, _i64, _vp, _vp]),
    "pcc_nn1_workspace_bytes": (_i64, [_i, _i, _i]),
    "pcc_nn1_f32": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "pcc_chamfer_workspace_bytes": (_i64, [_i, _i, _i]),
    "pcc_chamfer_fwd_f32": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pcc_chamfer_bwd_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
}



class PccMlpLayer(ctypes.Structure):
    """struct PccMlpLayer of include/pcc_b200.h"""
    _fields_ = [("packed_w", _vp), ("cin", _i), ("cout", _i), ("relu", _i), ("w_f32", _vp), ("b_f32", _vp)]


class PccMlpInput(ctypes.Structure):
    """struct PccMlpInput of include/pcc_b200.h"""
    _fields_ = [("ptr", _vp), ("dtype", _i), ("channels", _i), ("ld", _i64), ("row_div", _i)]


SIGNATURES.update({
    "pcc_debug_mlp_timing": (None, [_vp]),
    "pcc_debug_ws_timing": (None, [_vp]),
    "pcc_normalize_f32": (_i, [_vp, _i, _i, _f, _vp, _vp, _vp, _vp, _vp]),
    "pcc_assemble_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _vp, _vp]),
    "pcc_mlp_chain": (_i, [ctypes.POINTER(PccMlpInput), _i, _i64, ctypes.POINTER(PccMlpLayer), _i, _i, _vp, _i, _vp]),
    "pcc_mlp_packed_bytes": (_i64, [_i, _i]),
    "pcc_mlp_pack_weights_f32": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "pcc_mlp_chain_f32": (_i, [_vp, _i64, _i, ctypes.POINTER(PccMlpLayer), _i, _i, _vp, _vp]),
    "pcc_eval_metrics_f32": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    "pcc_pn_tail_bf16": (_i, [_vp, _i64, _i64, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "pcc_linear_bf16": (_i, [_vp, _i64, _i, _i64, _vp, _i64, _vp, _i, _i, _i, _vp, _i64, _vp]),
    "pcc_linear_small_f32": (_i, [_vp, _i, _i, _i64, _vp, _i64, _vp, _i, _i, _vp, _i64, _vp]),
    "pcc_fold_first_bf16": (_i, [_vp, _i, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _vp, _i64, _vp]),
    "pcc_gather_concat_bf16": (_i, [_vp, _i, _vp, _vp, _i, _i, _i64, _i, _vp, _vp, _i, _vp]),
    "pcc_knn_grid_workspace_bytes": (_i64, [_i, _i]),
    "pcc_knn_grid_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _f, _vp, _vp]),
    "pcc_pointnet_fused_bf16": (_i, [_vp, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "pcc_knn_patch_u8": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "pcc_linear_train_bf16": (_i, [_vp, _i64, _i, _i64, _vp, _i64, _vp, _i, _i, _vp, _i64, _i, _vp, _i64, _vp]),
    "pcc_groupmax_fwd_bf16": (_i, [_vp, _i64, _i, _i64, _i, _vp, _vp, _vp]),
    "pcc_groupmax_bwd_bf16": (_i, [_vp, _vp, _vp, _i64, _i, _i, _vp, _i64, _vp]),
    "pcc_wgrad_bf16": (_i, [_vp, _i64, _i, _vp, _i64, _i, _i64, _vp, _i64, _vp, _vp]),
    "pcc_sa_chain_indexed": (_i, [_vp, _vp, _i64, _i, ctypes.POINTER(PccMlpLayer), _i, _vp, _i, _vp]),
    "pcc_sa_chain_indexed_bwd": (_i, [_vp, _vp, _i64, _i] + [_vp] * 7 + [_i, _i64] + [_vp] * 7),
    "pcc_normals_pca_f32": (_i, [_vp, _i64, _i, _vp, _vp]),
    "pcc_p2plane_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "pcc_uc_f32": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "pcc_pmf_to_cdf_u16": (_i, [_vp, _i64, _i, _vp, _vp]),
    "pcc_cdf_to_u16": (_i, [_vp, _i64, _i, _vp, _vp]),
    "pcc_range_encode_u16": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp]),
    "pcc_range_decode_u16": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "pcc_quantise_latent_f32": (_i, [_vp, _i64, _i, _i, _f, _vp, _vp, _vp, _vp]),
    "pcc_octree_max_bits": (_i, [_i]),
    "pcc_octree_encode_f32": (_i, [_vp, _i, _i, _i, ctypes.c_double, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pcc_octree_decode_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
})

_lib = None


def load():
    """Load libpcc_b200.so and declare every entry point.  Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C point-cloud-compression_b200` "
                "(or __graft_entry__.build()); pcc_b200 has no CPU fallback")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().pcc_last_error_string().decode("utf-8", "replace")
        if rc < 0:
            raise ValueError(f"{what}: {msg} (code {rc})")
        raise RuntimeError(f"{what}: CUDA error {rc}: {msg}")
