"""Streamed GEMM layer (pcc_linear_bf16) against the library GEMM at the PPPF / inv_pool shapes."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
from pcc_b200 import mlp_ops
from tools.bench_ops import timeit

for M, K, N, group in [(2048, 1024, 16384, 0), (524288, 192, 128, 0), (524288, 128, 128, 0), (524288, 128, 256, 64),
                       (262144, 320, 256, 0), (262144, 256, 256, 0), (262144, 256, 512, 0), (262144, 512, 1024, 128),
                       (262144, 512, 1024, 0), (65536, 1024, 1024, 0)]:
    x = (torch.rand(M, K, device="cuda") - 0.5).to(torch.bfloat16)
    w = (torch.rand(N, K, device="cuda") - 0.5) / K ** 0.5
    b = torch.rand(N, device="cuda")
    t_ws, _ = timeit(lambda: mlp_ops.linear(x, w, b, True, group))
    def lib():
        y = torch.relu(torch.addmm(b.bfloat16(), x, w.bfloat16().t()))   # comparison arm only: plain torch bf16 GEMM
        return y.view(M // group, group, N).max(dim=1)[0] if group else y
    t_lib, _ = timeit(lib)
    fl = 2.0 * M * K * N
    print(f"M={M:7d} K={K:5d} N={N:6d} group={group:4d}: streamed {t_ws*1e3:8.1f} us ({fl/t_ws/1e9:7.1f} TFLOP/s)   "
          f"library {t_lib*1e3:8.1f} us ({fl/t_lib/1e9:7.1f} TFLOP/s)", flush=True)
