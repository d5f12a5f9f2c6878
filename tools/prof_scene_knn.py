"""cfg5 scene kNN (7812 queries x 1M points, K = 256) once warm, once more for an ncu capture of knn_warp_kernel."""
import sys, os, torch
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "point-cloud-compression_b200")]
import pcc_b200
from pcc_b200 import ops
g = torch.Generator(device="cuda").manual_seed(5)
sc = torch.rand(1, 1_000_000, 3, device="cuda", generator=g)
q = sc[:, torch.randperm(1_000_000, device="cuda", generator=g)[:7812]].contiguous()
for _ in range(2):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); ops.knn(q, sc, 256, True, True); e1.record(); torch.cuda.synchronize()
    print("knn 7812 x 1M K=256: %.3f ms" % e0.elapsed_time(e1), flush=True)
