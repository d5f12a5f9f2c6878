"""Mirror of the reference's own operator wrapper PointnetPPOps (/root/reference/pointnet_sa_module.py:8-34)."""
from . import pytorch3d_compat as p3d


class PointnetPPOps:
    @staticmethod
    def furthest_point_sample(xyz, npoint):
        _, idx = p3d.sample_farthest_points(xyz, K=npoint)
        return idx

    @staticmethod
    def ball_query(radius, nsample, xyz, new_xyz):
        return p3d.ball_query(new_xyz, xyz, K=nsample, radius=radius, return_nn=False)

    @staticmethod
    def group_points(features, idx):
        if hasattr(idx, "idx"):
            idx = idx.idx
        return p3d.knn_gather(features, idx.clamp(min=0))

    @staticmethod
    def knn_point(k, xyz, new_xyz):
        r = p3d.knn_points(new_xyz, xyz, K=k, return_nn=False)
        return r.dists, r.idx
