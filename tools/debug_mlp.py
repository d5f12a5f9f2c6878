"""Bring-up aid for the tcgen05 chain kernel: prints error statistics for a few tiny cases."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200"), os.path.join(ROOT, "tests")]
from pcc_b200 import mlp_ops
from test_gpu_mlp import make_layers, ref_chain

for dims, relu, rows, group in [([16, 32], [False], 128, 0), ([16, 128], [False], 128, 0), ([64, 128], [False], 128, 0),
                                ([16, 32, 64], [True, False], 128, 0), ([3, 32, 64, 128], [True, True, True], 256, 16)]:
    layers = make_layers(dims, relu, seed=sum(dims))
    x = (torch.rand(rows, dims[0], device="cuda") - 0.5) * 2
    y = mlp_ops.fused_chain(x, layers, group)
    torch.cuda.synchronize()
    ym = ref_chain(x, layers, group, True)
    print(dims, "rows", rows, "group", group, "max|y|", ym.abs().max().item(), "max err", (y - ym).abs().max().item(),
          "nan", torch.isnan(y).sum().item(), flush=True)
    if (y - ym).abs().max().item() > 1e-2:
        print(" y[0,:8]  ", y[0, :8].tolist())
        print(" ref[0,:8]", ym[0, :8].tolist())
        print(" y[1,:8]  ", y[1, :8].tolist())
        print(" ref[1,:8]", ym[1, :8].tolist())
        bad = ((y - ym).abs() > 1e-2)
        print(" bad rows:", bad.any(1).nonzero().flatten()[:16].tolist(), "bad cols:", bad.any(0).nonzero().flatten()[:16].tolist())
