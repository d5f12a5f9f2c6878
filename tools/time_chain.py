"""Per-phase clock64 ticks of the fused MLP chain (CTA 0, first tiles) -- bring-up aid."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200"), os.path.join(ROOT, "tests")]
from pcc_b200 import mlp_ops, _lib
from test_gpu_mlp import make_layers
lib = _lib.load()
dbg = lib.pcc_debug_mlp_timing
BS = 2048
cases = {
    "sa": (lambda: (torch.rand(BS * 256 * 16, 3, device="cuda") - 0.5), make_layers([3, 32, 64, 128], [True] * 3, 1), 16, torch.bfloat16),
    "pna": (lambda: [((torch.rand(BS * 256, 128, device="cuda") - 0.5).bfloat16(), 1), (torch.rand(BS * 256, 3, device="cuda"), 1)],
            make_layers([131, 128, 256], [True, True], 2), 0, torch.bfloat16),
}
for name, (mk, layers, group, od) in cases.items():
    x = mk()
    for _ in range(2):
        mlp_ops.fused_chain(x, layers, group, od)
    buf = torch.zeros(256, dtype=torch.int64, device="cuda")
    dbg(buf.data_ptr())
    mlp_ops.fused_chain(x, layers, group, od)
    torch.cuda.synchronize()
    dbg(None)
    t = buf.cpu().tolist()
    n = max(i for i, v in enumerate(t) if v) + 1
    f32 = 1 if (len(layers) >= 2 and layers[0][0].shape[1] <= 7) else 0
    L = len(layers)
    per = 4 + 7 * (L - f32)
    print(name, "ticks per tile:", per)
    for tile in range(2, min(8, n // per)):
        seg = t[tile * per:(tile + 1) * per + 1]
        d = [seg[i + 1] - seg[i] for i in range(len(seg) - 1)]
        labels = ["stage", "cpwait", "prefetch"] + sum([[f"L{l} pre", f"L{l} mma", f"L{l} commit", f"L{l} wait", f"L{l} cpissue", f"L{l} epi", f"L{l} sync"] for l in range(f32, L)], []) + ["loop"]
        print(f" tile {tile}: total {seg[-1] - seg[0]:6d} | " + " ".join(f"{a}={b}" for a, b in zip(labels, d)))
