"""Put the B200 ops in front of the unmodified reference.

install() registers `pytorch3d`, `pytorch3d.ops`, `pytorch3d.ops.knn` and `pytorch3d.loss` modules backed by
pcc_b200 (the reference imports those names at module import time) and, when the reference's pn_kit /
pppe_pcd_ae / pointnet_sa_module are (or later get) imported, rebinds their FPS / gather helpers.
Use:  import pcc_b200; pcc_b200.install(); import pn_kit, AE, ...      
"""
import sys
import types

from . import octree_ops, pn_kit_ops, pointnet_ops, pytorch3d_compat as p3d, torchac_compat


def _module(name):
    m = types.ModuleType(name)
    m.__pcc_b200__ = True
    sys.modules[name] = m
    return m


def install(patch_loaded=True):
    root, ops_m, knn_m, loss_m = (_module(n) for n in ("pytorch3d", "pytorch3d.ops", "pytorch3d.ops.knn",
                                                       "pytorch3d.loss"))
    root.ops, root.loss, ops_m.knn = ops_m, loss_m, knn_m
    for m in (ops_m, knn_m):
        m._KNN = p3d._KNN
        m.knn_points = p3d.knn_points
        m.knn_gather = p3d.knn_gather
    ops_m.ball_query = p3d.ball_query
    ops_m.sample_farthest_points = p3d.sample_farthest_points
    loss_m.chamfer_distance = p3d.chamfer_distance
    tac = _module("torchac")                                      # compress.py / decompress.py: `import torchac`
    tac.encode_float_cdf, tac.decode_float_cdf = torchac_compat.encode_float_cdf, torchac_compat.decode_float_cdf
    if patch_loaded:
        patch_reference_modules()


def patch_reference_modules():
    """Rebind names the reference modules captured at import (pn_kit.py:309-360, pppe_pcd_ae.py:7,551)."""
    pn = sys.modules.get("pn_kit")
    if pn is not None:
        pn.farthest_point_sample_batch = pn_kit_ops.farthest_point_sample_batch
        pn.index_points = pn_kit_ops.index_points
        pn.knn_points, pn.knn_gather = p3d.knn_points, p3d.knn_gather
        pn.encode_sampled_np = octree_ops.encode_sampled_np          # pn_kit.py:380-401
        pn.decode_sampled_np = octree_ops.decode_sampled_np          # pn_kit.py:424-431
    on = sys.modules.get("octree_np")
    if on is not None:
        on.encode, on.decode = octree_ops.encode, octree_ops.decode  # octree_np.py:10-45, 47-112
    pp = sys.modules.get("pppe_pcd_ae")
    if pp is not None:
        pp.farthest_point_sample_batch = pn_kit_ops.farthest_point_sample_batch
        pp.index_points = pn_kit_ops.index_points
        pp.knn_points, pp.chamfer_distance = p3d.knn_points, p3d.chamfer_distance
    sa = sys.modules.get("pointnet_sa_module")
    if sa is not None:
        sa.PointnetPPOps = pointnet_ops.PointnetPPOps
        sa.sample_farthest_points, sa.knn_points = p3d.sample_farthest_points, p3d.knn_points
        sa.knn_gather, sa.ball_query = p3d.knn_gather, p3d.ball_query
    for name in ("AE", "PPPF_AE"):
        m = sys.modules.get(name)
        if m is not None:
            m.chamfer_distance = p3d.chamfer_distance
