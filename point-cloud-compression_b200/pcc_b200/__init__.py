"""pcc_b200 -- B200-native (sm_100a) hot path of the IPDAE-style point-cloud codec rhmes/point-cloud-compression.

Host side of libpcc_b200.so (C ABI in include/pcc_b200.h).  Public surface = the reference's own names:
  pcc_b200.pn_kit_ops.{farthest_point_sample_batch, index_points}         (pn_kit.py:309-360)
  pcc_b200.octree_ops.{encode_sampled_np, decode_sampled_np, encode, decode}  (pn_kit.py:380-431, octree_np.py:10-112)
  pcc_b200.pointnet_ops.PointnetPPOps                                     (pointnet_sa_module.py:8-34)
  pcc_b200.pytorch3d_compat.{knn_points, knn_gather, ball_query, sample_farthest_points, chamfer_distance}
  pcc_b200.install()  -> registers the pytorch3d.* module names and patches loaded reference modules.
"""
from . import _lib, bodies, octree_ops, ops, pn_kit_ops, pointnet_ops, pytorch3d_compat, torchac_compat  # noqa: F401
from .install import install, patch_reference_forwards, patch_reference_modules, uninstall  # noqa: F401
from .pn_kit_ops import farthest_point_sample_batch, index_points  # noqa: F401
from .pointnet_ops import PointnetPPOps  # noqa: F401
from .pytorch3d_compat import (ball_query, chamfer_distance, knn_gather, knn_points,  # noqa: F401
                               sample_farthest_points)

__version__ = "0.1.0"
