"""Chamfer (distance-only and keyed) on the two regimes: real reconstructions of the headline workload (clumpy) and noisy copies
(close clouds, cfg4's regime); PCC_CHAMFER_GRID=G overrides the grid resolution."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
import pcc_b200
from pcc_b200.codec import PatchCodec
from pcc_b200.modules import AE
from tools import synth
from tools.bench_ops import timeit
ae = AE(256, 128, 16, 7)
ae.load_state_dict(synth.seeded_state_dict(synth.ae_shapes(128, 16, 7), 11))
ae = ae.cuda().eval()
codec = PatchCodec(ae, centre_mode="coded")
xyz = torch.from_numpy(synth.modelnet_like(32, 8192, seed=1000)).cuda()
start = torch.zeros(32, dtype=torch.int64, device="cuda")
c = codec.compress(xyz, start)
rec = codec.decompress(c["latent_q"], c["centres"], 8192, c["center"], c["longest"])
noisy = torch.from_numpy(synth.decompressed_like(xyz.cpu().numpy(), seed=8)).cuda()
for name, y in (("clumpy", rec), ("close", noisy)):
    for idx in (False, True):
        b, m = timeit(lambda: pcc_b200.ops.chamfer_forward(y, xyz, want_idx=idx), iters=20)
        print(f"G={os.environ.get('PCC_CHAMFER_GRID', 'default')} {name:6s} want_idx={idx}: best {b * 1e3:.1f} us", flush=True)
