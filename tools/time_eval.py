"""Device time of the eval stage (Chamfer + D1) on real reconstructions of the headline workload, per stage of roundtrip."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
import pcc_b200
from pcc_b200.codec import PatchCodec
from pcc_b200.modules import AE
from tools import synth

ae = AE(256, 128, 16, 7)
ae.load_state_dict(synth.seeded_state_dict(synth.ae_shapes(128, 16, 7), 11))
ae = ae.cuda().eval()
codec = PatchCodec(ae, centre_mode="coded")
xyz = torch.from_numpy(synth.modelnet_like(32, 8192, seed=1000)).cuda()
start = torch.zeros(32, dtype=torch.int64, device="cuda")
c = codec.compress(xyz, start)
rec = codec.decompress(c["latent_q"], c["centres"], 8192, c["center"], c["longest"])

def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[0] * 1e3

print(f"compress   {t(lambda: codec.compress(xyz, start)):8.1f} us")
print(f"decompress {t(lambda: codec.decompress(c['latent_q'], c['centres'], 8192, c['center'], c['longest'])):8.1f} us")
print(f"evaluate   {t(lambda: codec.evaluate(rec, xyz, c['bbox'])):8.1f} us")
print(f"chamfer    {t(lambda: pcc_b200.ops.chamfer_forward(rec, xyz, want_idx=False)):8.1f} us   (PCC_CHAMFER_PATH={os.environ.get('PCC_CHAMFER_PATH')})")
print(f"roundtrip  {t(lambda: codec.roundtrip(xyz, start)):8.1f} us")
