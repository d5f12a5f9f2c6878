"""Training step of the IPDAE patch auto-encoder on the B200 ops: the body of train.py:148-247 for one batch.

The data-parallel geometric stages (normalise, FPS, kNN patching, in-patch kNN, Chamfer forward / backward) are the pcc
kernels, and so are the network bodies: every contraction of the forward and backward pass runs on the streamed tcgen05 GEMM,
the MN-major weight-gradient kernel and the pooling kernels (train_ops.py: bf16 operands, fp32 accumulation and weight
gradients, fp32 master weights -- the counterpart of the reference's autocast path, train.py:114,154).  Trainer(kernels=False)
runs the same bodies as plain fp32 torch ops (library GEMMs): the arithmetic the kernel path is tested against.  Multi-GPU: one process per GPU, whole clouds per rank, gradients all-reduced by DistributedDataParallel
over NCCL -- the only collective of the path.  The octree centre coder stays on the reference path; the step applies its
quantisation rule on the device (see codec.py).
"""
import contextlib
import math

import os

import torch

from . import ops
from .modules import AE, ConditionalProbabilityModel
from .pytorch3d_compat import chamfer_distance


def estimate_bits_from_pmf(pmf, sym):
    """pn_kit.estimate_bits_from_pmf (/root/reference/pn_kit.py:439-450)."""
    L = pmf.shape[-1]
    p = torch.gather(pmf.reshape(-1, L), dim=1, index=sym.reshape(-1, 1))
    return torch.sum(-torch.log2(p.clamp(min=1e-3)))


@contextlib.contextmanager
def _matmul_tf32(enabled):
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = bool(enabled) or old
    try:
        yield
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


class Trainer:
    def __init__(self, K=256, k=128, d=16, L=7, N0=1024, alpha=2, lr=0.0005, lamda=1e-6, rate_loss_enable_step=40000,
                 centre_depth=6, device="cuda", ddp=False, state_dict=None, tf32=True, amp=False, kernels=True):
        self.K, self.k, self.d, self.L, self.N0, self.alpha = K, k, d, L, N0, alpha
        # The reference's network bodies are 1x1 Conv2d layers, which PyTorch runs through cuDNN with TF32 enabled by
        # default (torch.backends.cudnn.allow_tf32); the addmm form used here gets the same arithmetic only when the
        # matmul flag is switched on as well.  fp32 storage and accumulation, 10-bit operand mantissas.
        # The flag is set only around this trainer's own forward / backward (a context in step()), never process-wide: other
        # code in the process -- e.g. a probability model whose PMFs must be reproduced bit for bit -- keeps its setting.
        self.kernels = bool(kernels)
        self.tf32 = bool(tf32) and not self.kernels      # tf32 / amp only concern the library-GEMM bodies (kernels=False)
        amp = bool(amp) and not self.kernels
        # amp=True: the network bodies run under bf16 autocast -- the counterpart of the reference's fp16 autocast + GradScaler
        # path (train.py:114,154-160, taken when --device is the string 'cuda'); bf16 needs no loss scaling.
        self.amp = amp
        self.lamda, self.rate_loss_enable_step, self.centre_depth = lamda, rate_loss_enable_step, centre_depth
        self.ae = AE(K, k, d, L).to(device)
        if state_dict is not None:
            self.ae.load_state_dict(state_dict)
        self.prob = ConditionalProbabilityModel(L, d).to(device)
        name = "forward_train" if self.kernels else "forward_train_fp32"
        self.ae_fwd, self.prob_fwd = getattr(self.ae, name), getattr(self.prob, name)
        if ddp:  # gradients of both models are all-reduced over NCCL, bucketed and overlapped with the backward pass
            from torch.nn.parallel import DistributedDataParallel as DDP

            class _Train(torch.nn.Module):
                def __init__(self, net):
                    super().__init__()
                    self.net = net

                def forward(self, x):
                    return getattr(self.net, name)(x)

            self._ddp_ae, self._ddp_prob = DDP(_Train(self.ae)), DDP(_Train(self.prob))
            self.ae_fwd, self.prob_fwd = self._ddp_ae, self._ddp_prob
        # capturable: the step counters live on the device, so the whole step (forward, backward, Adam) can be replayed as one
        # CUDA graph (step_graphed); the update arithmetic is the reference's torch.optim.Adam (train.py:132-135)
        # fused: torch's single multi-tensor CUDA kernel for the same update (the default "foreach" form is ~100 small launches per
        # step: 0.95 ms of the 6.9 ms step); PCC_ADAM_FOREACH=1 keeps the default form
        on_cuda = str(device).startswith("cuda")
        self.optimizer = torch.optim.Adam(list(self.ae.parameters()) + list(self.prob.parameters()), lr=lr, capturable=on_cuda,
                                          fused=on_cuda and not os.environ.get("PCC_ADAM_FOREACH"))
        self.global_step = 0
        self.ddp = bool(ddp)
        self._graph = None

    def step(self, batch_x, start_idx=None):
        """One optimisation step on batch_x [B,N,3] (device).  Returns dict(loss, chamfer, fbpp)."""
        self.ae.train()
        self.prob.train()
        B, N, _ = batch_x.shape
        S = N * self.alpha // self.K
        # train.py:164 -- pn_kit.normalize takes the bounding box of cloud 0 for the whole batch (appendix B-5)
        _, center, longest, _ = ops.normalize(batch_x[:1])
        x = (batch_x - center.view(1, 1, 3)) * (1 - 0.01) / longest.view(1, 1, 1) + 0.5
        self.optimizer.zero_grad(set_to_none=True)
        if start_idx is None:                                                     # pn_kit.py:321 (CPU RNG)
            start_idx = torch.randint(0, N, (B,), dtype=torch.long).to(x.device)
        cube = 1.0 / max(1.0, math.pow(2.0, min(self.centre_depth, 30)))
        _, rec_centres = ops.fps(x, S, start_idx, 1e10, return_xyz=True, quant_cube=cube)          # train.py:171-179
        scale = (N / self.N0) ** (1 / 3)
        _, _, patches = ops.knn(rec_centres, x, self.K, return_nn=True, centre_sub=True, nn_scale=scale,
                                nn_only=True)                                     # train.py:185-192
        with _matmul_tf32(self.tf32):
            return self._network_step(batch_x, x, rec_centres, patches, scale, B, N, S)

    def step_graphed(self, batch_x, start_idx):
        """step() replayed as ONE captured CUDA graph (forward + Chamfer + backward + Adam: ~300 kernel launches whose host-side
        issue time exceeds their device time).  Needs a fixed batch shape and an explicit start_idx (the CPU RNG draw cannot be
        captured); the first call warms up with three eager steps and captures, a shape change or the rate-loss switch
        (train.py:211-214) re-captures.  Returns the graph's static result tensors (overwritten by the next call).  Single
        process only: under DistributedDataParallel the eager step runs (its gradient hooks are not captured)."""
        if self.ddp:
            return self.step(batch_x, start_idx)
        lam_on = self.global_step >= self.rate_loss_enable_step
        key = (tuple(batch_x.shape), lam_on)
        if self._graph is None or self._graph["key"] != key:
            from . import mlp_ops
            sx, ss = batch_x.clone(), start_idx.clone()
            side = torch.cuda.Stream(batch_x.device)
            side.wait_stream(torch.cuda.current_stream(batch_x.device))
            with torch.cuda.stream(side):
                for _ in range(3):
                    self.step(sx, ss)
            torch.cuda.current_stream(batch_x.device).wait_stream(side)
            self.optimizer.zero_grad(set_to_none=True)
            mlp_ops.invalidate_weight_caches()   # the bf16 weight copies must be re-made INSIDE the graph on every replay
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self.step(sx, ss)
            self._graph = dict(key=key, graph=g, x=sx, start=ss, out=out)
            return out
        gr = self._graph
        gr["x"].copy_(batch_x, non_blocking=True)
        gr["start"].copy_(start_idx, non_blocking=True)
        gr["graph"].replay()
        self.global_step += 1
        return gr["out"]

    def _network_step(self, batch_x, x, rec_centres, patches, scale, B, N, S):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.amp):
            patches_pred, _, latent_q = self.ae_fwd(patches.view(B * S, self.K, 3))   # train.py:193
            pmf = self.prob_fwd(rec_centres)                                      # train.py:197
        patches_pred, latent_q, pmf = patches_pred.float(), latent_q.float(), pmf.float()
        patches_pred = patches_pred / scale                                       # train.py:194
        sym = (latent_q.view(B, S, self.d) + self.L // 2).long().clamp(0, self.L - 1)
        feature_bits = estimate_bits_from_pmf(pmf, sym) / (B * N)                 # train.py:201
        fbpp = feature_bits / (B * N)                                             # train.py:205 (divided twice, appendix B-7)
        pc_pred = (patches_pred.view(B, S, -1, 3) + rec_centres.view(B, S, 1, 3)).reshape(B, -1, 3)  # train.py:207-209
        cham, _ = chamfer_distance(pc_pred, x)                                    # AE.get_loss, AE.py:67
        lam = 0.0 if self.global_step < self.rate_loss_enable_step else self.lamda
        loss = cham + lam * fbpp                                                  # AE.py:68-69
        loss.backward()                                                           # train.py:221
        self.optimizer.step()
        if self.kernels:   # the fused optimiser kernel does not bump the parameters' version counters: cached bf16 copies are stale
            from . import mlp_ops
            mlp_ops.invalidate_weight_caches()
        self.global_step += 1
        return dict(loss=loss.detach(), chamfer=cham.detach(), fbpp=fbpp.detach())
