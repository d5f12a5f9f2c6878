"""Put the B200 ops in front of the unmodified reference.

install() registers `pytorch3d`, `pytorch3d.ops`, `pytorch3d.ops.knn` and `pytorch3d.loss` modules backed by
pcc_b200 (the reference imports those names at module import time) and, when the reference's pn_kit /
pppe_pcd_ae / pointnet_sa_module / AE / PPPF_AE are (or later get) imported, rebinds their FPS / gather helpers AND swaps the
`forward` of their network classes for the fused bodies (bodies.py), which read the parameters of the reference's own module
instances by the reference's attribute names.  The swapped forward keeps the reference's signature and layouts; it hands the
call back to the reference's original forward whenever the call must stay differentiable (autograd on and trainable
parameters: train.py's loop), or BatchNorm is in training mode (batch statistics), or the input is not on the GPU.
Use:  import pcc_b200; pcc_b200.install(); import pn_kit, AE, ...      
"""
import functools
import importlib.abc
import importlib.util
import sys
import types

import torch

from . import bodies, octree_ops, pn_kit_ops, pointnet_ops, pytorch3d_compat as p3d, torchac_compat


def _module(name):
    m = types.ModuleType(name)
    m.__pcc_b200__ = True
    sys.modules[name] = m
    return m


_REFERENCE_MODULES = ("pn_kit", "octree_np", "pppe_pcd_ae", "pointnet_sa_module", "AE", "PPPF_AE")


class _PatchOnImport(importlib.abc.MetaPathFinder):
    """Reference modules imported AFTER install() (compress.py imports pn_kit and AE at its top, after the launcher has
    run) are patched the moment their import finishes."""

    def __init__(self):
        self._busy = False

    def find_spec(self, name, path=None, target=None):
        if name not in _REFERENCE_MODULES or self._busy:
            return None
        self._busy = True
        try:
            spec = importlib.util.find_spec(name)
        except (ImportError, ValueError):
            spec = None
        finally:
            self._busy = False
        if spec is None or spec.loader is None or not hasattr(spec.loader, "exec_module"):
            return None
        loader, run = spec.loader, spec.loader.exec_module

        def exec_module(module):
            run(module)
            patch_reference_modules()

        loader.exec_module = exec_module
        return spec


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
    if not any(isinstance(f, _PatchOnImport) for f in sys.meta_path):
        sys.meta_path.insert(0, _PatchOnImport())
    if patch_loaded:
        patch_reference_modules()


def uninstall():
    """Undo install(): drop the import hook, restore the reference's own forward bodies, remove the shim modules.  (Helper
    names already rebound inside loaded reference modules stay rebound; tests that need the pristine reference re-import it.)"""
    sys.meta_path[:] = [f for f in sys.meta_path if not isinstance(f, _PatchOnImport)]
    unpatch_reference_forwards()
    for name in ("pytorch3d", "pytorch3d.ops", "pytorch3d.ops.knn", "pytorch3d.loss", "torchac"):
        if getattr(sys.modules.get(name), "__pcc_b200__", False):
            del sys.modules[name]


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
    patch_reference_forwards()


# ---- forward bodies of the reference's network classes --------------------------------------------------------------------
def _fused_ok(mod, x):
    return (isinstance(x, torch.Tensor) and x.is_cuda and not bodies.training_pass(mod) and not bodies.has_train_mode_bn(mod))


def _swap_forward(cls, fused):
    """cls.forward := fused(self, *args) when the call is an inference call on the GPU, the original forward otherwise."""
    orig = cls.forward
    if getattr(orig, "__pcc_b200__", False):
        return

    @functools.wraps(orig)
    def forward(self, x, *args, **kwargs):
        if _fused_ok(self, x):
            with torch.no_grad():
                return fused(self, x, *args, **kwargs)
        return orig(self, x, *args, **kwargs)

    forward.__pcc_b200__ = True
    forward.__pcc_original__ = orig
    cls.forward = forward


def _swap_linear_stack(cls, attr):
    """Instances of `cls` get their `attr` nn.Sequential (Linear + ReLU stack) re-classed after construction, so that calling it
    on its own -- decompress.py:96 `ae.inv_pool(latent_quantized)` -- runs the streamed GEMM kernels.  A subclass of nn.Sequential
    with the same children keeps the state_dict keys."""
    init = cls.__init__
    if getattr(init, "__pcc_b200__", False):
        return

    class FusedLinearStack(torch.nn.Sequential):
        def forward(self, x):
            if _fused_ok(self, x):
                with torch.no_grad():
                    return bodies.linear_stack_forward(self, x)
            return super().forward(x)

    @functools.wraps(init)
    def __init__(self, *args, **kwargs):
        init(self, *args, **kwargs)
        seq = getattr(self, attr, None)
        if type(seq) is torch.nn.Sequential:
            seq.__class__ = FusedLinearStack

    __init__.__pcc_b200__ = True
    __init__.__pcc_original__ = init
    cls.__init__ = __init__


def _sa_forward(mod, xyz):
    """pn_kit.SetAbstraction.forward: xyz [B, 3, N] -> (new_xyz [B, 3, S], new_points [B, D', S])   (pn_kit.py:164-211)."""
    new_xyz, feat = bodies.sa_points(mod, xyz.permute(0, 2, 1).contiguous())
    return new_xyz.permute(0, 2, 1), feat.permute(0, 2, 1)


def _pointnet_forward(mod, points):
    """pn_kit.PointNet.forward: points [B, C, N] -> [B, D]   (pn_kit.py:124-144)."""
    return bodies.pointnet_points(mod, points.permute(0, 2, 1).contiguous())


def _mlp_forward(mod, points):
    """pn_kit.MLP.forward: points [B, C, N] -> [B, D, N]   (pn_kit.py:289-305)."""
    return bodies.mlp_points(mod, points.permute(0, 2, 1).contiguous()).permute(0, 2, 1)


def patch_reference_forwards():
    """Swap the forward bodies of the reference's network classes that are loaded (idempotent).
    pn_kit.{SetAbstraction, PointNet, MLP} pn_kit.py:124-211,289-305; AE.{AE, ConditionalProbabilityModel} AE.py:34-55,107-123;
    pointnet_sa_module.PointnetSAModule pointnet_sa_module.py:58-93; PPPF_AE.{PointNetPP, FoldingNet, PPPF_AE,
    ConditionalProbabilityModel} PPPF_AE.py:39-46,91-109,128-150,203-228; pppe_pcd_ae.{PointNetSetAbstraction,
    PointNetSetAbstractionMSG, PointNet2EncoderFull} pppe_pcd_ae.py:588-618,627-633,672-690."""
    table = (("pn_kit", (("SetAbstraction", _sa_forward), ("PointNet", _pointnet_forward), ("MLP", _mlp_forward))),
             ("AE", (("AE", bodies.ae_forward), ("ConditionalProbabilityModel", bodies.prob_forward))),
             ("pointnet_sa_module", (("PointnetSAModule", bodies.sa_module_forward),)),
             ("PPPF_AE", (("PointNetPP", bodies.pointnetpp_forward), ("FoldingNet", bodies.folding_forward),
                          ("PPPF_AE", bodies.pppf_forward), ("ConditionalProbabilityModel", bodies.pppf_prob_forward))),
             ("pppe_pcd_ae", (("PointNetSetAbstraction", bodies.pppe_sa_forward), ("PointNetSetAbstractionMSG", bodies.pppe_sa_forward),
                              ("PointNet2EncoderFull", bodies.pppe_encoder_forward))))
    for mod_name, classes in table:
        m = sys.modules.get(mod_name)
        if m is None or getattr(m, "__pcc_b200__", False):
            continue
        for cls_name, fused in classes:
            cls = getattr(m, cls_name, None)       # a module that is still being imported does not have its classes yet
            if isinstance(cls, type):
                _swap_forward(cls, fused)
                if (mod_name, cls_name) == ("AE", "AE"):
                    _swap_linear_stack(cls, "inv_pool")               # AE.py:19-26, called alone at decompress.py:96


def unpatch_reference_forwards():
    """Restore the reference's own forward bodies (tests compare the two)."""
    for name, classes in (("pn_kit", ("SetAbstraction", "PointNet", "MLP")), ("AE", ("AE", "ConditionalProbabilityModel")),
                          ("pointnet_sa_module", ("PointnetSAModule",)),
                          ("PPPF_AE", ("PointNetPP", "FoldingNet", "PPPF_AE", "ConditionalProbabilityModel")),
                          ("pppe_pcd_ae", ("PointNetSetAbstraction", "PointNetSetAbstractionMSG", "PointNet2EncoderFull"))):
        m = sys.modules.get(name)
        for c in classes if m is not None else ():
            cls = getattr(m, c, None)
            orig = getattr(getattr(cls, "forward", None), "__pcc_original__", None)
            if orig is not None:
                cls.forward = orig
            init = getattr(getattr(cls, "__init__", None), "__pcc_original__", None)
            if init is not None:
                cls.__init__ = init
