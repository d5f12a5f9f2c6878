"""BASELINE cfg4: eval.py's Chamfer / D1-PSNR sweep over 10,000 synthetic 8192-point clouds, whole clouds sharded by rank
(python tools/eval_sweep.py, or torchrun --nproc-per-node N tools/eval_sweep.py).  The "decompressed" clouds are the originals
plus N(0, 2e-3) noise in a shuffled order; every chunk of 256 clouds is one evaluate() call.  Prints clouds/s
(device time, max over ranks) and the mean metrics gathered from all ranks -- the only collective of the path."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
from pcc_b200.codec import PatchCodec
from pcc_b200.modules import AE
from pcc_b200 import dist as pdist
from tools import synth

TOTAL, CHUNK = int(os.environ.get("CLOUDS", 10000)), 256
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
codec = PatchCodec(AE(256, 128, 16, 7).to(dev).eval())
base = torch.from_numpy(synth.modelnet_like(CHUNK, 8192, seed=500 + rank)).to(dev)
g = torch.Generator(device=dev).manual_seed(rank)
mine = (TOTAL + world - 1) // world                     # whole clouds per rank
chunks = (mine + CHUNK - 1) // CHUNK
mets = []
def chunk(c):
    x = base * (1.0 - 1e-4 * c)                         # a different cloud set per chunk
    y = (x + 2e-3 * torch.randn(x.shape, device=dev, generator=g))[:, torch.randperm(8192, device=dev, generator=g)]
    return codec.evaluate(y, x)
for c in range(2):
    chunk(c)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for c in range(chunks):
    mets.append(chunk(c))
m = torch.cat(mets)[:mine]
if world > 1:
    out = [torch.empty_like(m) for _ in range(world)]
    dist.all_gather(out, m)
    m = torch.cat(out)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"cfg4 eval sweep: {m.shape[0]} clouds on {world} GPU(s) in {ms.item():.1f} ms -> {m.shape[0] / ms.item() * 1e3:.0f} clouds/s "
          f"(incl. synthesising the noisy copies); mean Chamfer {m[:, 0].mean().item():.3e}, mean D1 PSNR {m[:, 1].mean().item():.2f} dB")
if world > 1:
    dist.destroy_process_group()
