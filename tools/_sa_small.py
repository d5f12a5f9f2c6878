import os, sys, torch
ROOT = "/root/repo"
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200"), os.path.join(ROOT, "tests")]
from pcc_b200 import mlp_ops
from test_gpu_mlp import make_layers, ref_chain
torch.manual_seed(0)
sa = make_layers([3, 32, 64, 128], [True, True, True], 1)
for rows in (128, 256, 1024):
    g = torch.rand(rows, 3, device="cuda") - 0.5
    try:
        y = mlp_ops.fused_chain(g, sa, group=16)
        torch.cuda.synchronize()
        ym = ref_chain(g, sa, 16, model_bf16=True)
        bad = ((y - ym).abs() > 2e-3 * ym.abs().max()) | ~torch.isfinite(y)
        print("flags", os.environ.get("PCC_SA_DBG"), "rows", rows, "bad frac", bad.float().mean().item(), flush=True)
        if bad.any():
            print(" bad per out-row:", bad.sum(1).tolist()[:32])
            print(" bad per channel quadrant:", bad.view(-1, 4, 32).sum((0, 2)).tolist())
            r = bad.nonzero()[0]
            print(" first bad", r.tolist(), y[r[0], r[1]].item(), ym[r[0], r[1]].item())
            print(" y row0[:8]", y[0, :8].tolist(), "\n ym row0[:8]", ym[0, :8].tolist())
    except Exception as e:
        print("flags", os.environ.get("PCC_SA_DBG"), "FAIL", str(e)[:100])
        break
