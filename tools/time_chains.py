"""Device time of the AE's three fused chains at the headline size (CUDA events, best of N)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200"), os.path.join(ROOT, "tests")]
from pcc_b200 import mlp_ops
from test_gpu_mlp import make_layers

BS = int(os.environ.get("BS", 2048))
sa = make_layers([3, 32, 64, 128], [True, True, True], 1)
pna = make_layers([131, 128, 256], [True, True], 2)
dec = make_layers([144, 128, 64, 32, 3], [True, True, True, False], 3)
g = torch.rand(BS * 256 * 16, 3, device="cuda") - 0.5
xyz = torch.rand(BS * 256, 3, device="cuda") - 0.5
lat = torch.randint(-3, 4, (BS, 16), device="cuda").float()
lin = (torch.rand(BS * 128, 128, device="cuda") - 0.5).bfloat16()
feat = mlp_ops.fused_chain(g, sa, group=16, out_dtype=torch.bfloat16)
cases = {
    "sa  3-32-64-128 max16": (lambda: mlp_ops.fused_chain(g, sa, group=16, out_dtype=torch.bfloat16), BS * 256 * 16 * 2 * (3 * 32 + 32 * 64 + 64 * 128)),
    "pnf 131-128-256": (lambda: mlp_ops.fused_chain([(feat, 1), (xyz, 1)], pna, out_dtype=torch.bfloat16), BS * 256 * 2 * (131 * 128 + 128 * 256)),
    "dec 144-128-64-32-3": (lambda: mlp_ops.fused_chain([(lin, 1), (lat, 128)], dec), BS * 128 * 2 * (144 * 128 + 128 * 64 + 64 * 32 + 32 * 3)),
}
tail = make_layers([256, 512, 16], [True, False], 4)
h256 = mlp_ops.fused_chain([(feat, 1), (xyz, 1)], pna, out_dtype=torch.bfloat16)
cases["pn tail 256-512-16 max256"] = (lambda: mlp_ops.pn_tail(h256, tail), BS * 256 * 2 * (256 * 512 + 512 * 16))



def _library_tail():   # comparison arm only (tools/): the same two layers as plain torch bf16 GEMMs + a separate max
    h = torch.relu(torch.addmm(tail[0][1].bfloat16(), h256, tail[0][0].bfloat16().t()))
    return torch.addmm(tail[1][1].bfloat16(), h, tail[1][0].bfloat16().t()).view(BS, 256, -1).max(dim=1)[0]


cases["pn tail (library GEMMs)"] = (_library_tail, BS * 256 * 2 * (256 * 512 + 512 * 16))
for name, (fn, flop) in cases.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"{name:24s} best {ts[0]*1e3:8.1f} us  median {ts[5]*1e3:8.1f} us  {flop/ts[0]/1e9:8.1f} TFLOP/s  (PCC_NO_WS={os.environ.get('PCC_NO_WS')})", flush=True)
