"""cfg4-shaped eval sweep (10,000 clouds x 8192 points, chunks of 256) through PatchCodec.evaluate_sweep with 1 / 2 / 3 streams."""
import os, sys, torch
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "point-cloud-compression_b200")]
from pcc_b200.codec import PatchCodec
from pcc_b200.modules import AE
from tools import synth
ae = AE(256, 128, 16, 7); ae.load_state_dict(synth.seeded_state_dict(synth.ae_shapes(128, 16, 7), 11))
codec = PatchCodec(ae.cuda().eval())
chunk, total = 256, 10000
base = torch.from_numpy(synth.modelnet_like(chunk, 8192, seed=500)).cuda()
noisy = torch.from_numpy(synth.decompressed_like(base.cpu().numpy(), seed=501)).cuda()
def chunks():
    for i in range(0, total, chunk):
        n = min(chunk, total - i); s = 1.0 - 1e-6 * (i // chunk)
        yield noisy[:n] * s, base[:n] * s
for ns in (1, 2, 3, 1, 2):
    codec.evaluate_sweep(chunks(), ns); torch.cuda.synchronize()
    t = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); r = codec.evaluate_sweep(chunks(), ns); e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    print(f"evaluate_sweep {total} clouds, {ns} stream(s): best {min(t):.2f} ms", flush=True)
