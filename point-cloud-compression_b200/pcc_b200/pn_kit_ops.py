"""Drop-ins for the reference's in-tree helpers (/root/reference/pn_kit.py:309-360), same names and signatures."""
import torch

from . import ops


def farthest_point_sample_batch(xyz, npoint):
    """pn_kit.farthest_point_sample_batch(xyz [B,N,3], npoint) -> int64 [B,npoint].

    The start index is drawn exactly as the reference does (pn_kit.py:321): torch.randint on the CPU global
    generator, then moved to the device -- so a seeded script sees the same draws and the same centres."""
    B, N, _ = xyz.shape
    farthest = torch.randint(0, N, (B,), dtype=torch.long).to(xyz.device)
    return ops.fps(xyz, npoint, farthest, 1e10)


def index_points(points, idx):
    """pn_kit.index_points(points [B,N,C], idx [B,S] or [B,S,K]) -> [B,S,C] / [B,S,K,C]."""
    return ops.gather(points, idx)
