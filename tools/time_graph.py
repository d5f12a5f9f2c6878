"""CUDA-graph replay of the single-cloud round trip (cfg1 latency): capture once, replay with a copy-in."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
from pcc_b200.codec import PatchCodec
from pcc_b200.modules import AE
from tools import synth
from tools.bench_ops import timeit

B = int(os.environ.get("B", 1))
ae = AE(256, 128, 16, 7)
ae.load_state_dict(synth.seeded_state_dict(synth.ae_shapes(128, 16, 7), 11))
codec = PatchCodec(ae.cuda().eval(), centre_mode="coded")
x = torch.from_numpy(synth.modelnet_like(B, 8192, seed=1)).cuda()
start = torch.zeros(B, dtype=torch.int64, device="cuda")
ref = codec.roundtrip(x, start)
print("eager", timeit(lambda: codec.roundtrip(x, start)))
run = codec.graphed_roundtrip(B, 8192)
out = run(x, start)
torch.cuda.synchronize()
assert torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1]) and torch.equal(out[2], ref[2]) and torch.equal(out[3], ref[3])
x2 = torch.from_numpy(synth.modelnet_like(B, 8192, seed=2)).cuda()
ref2 = codec.roundtrip(x2, start)
out2 = run(x2, start)
torch.cuda.synchronize()
assert torch.equal(out2[2], ref2[2]) and torch.equal(out2[3], ref2[3])
print("graph", timeit(lambda: run(x, start)))
