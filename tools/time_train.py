import os, sys, torch
sys.path[:0] = ["/root/repo", "/root/repo/point-cloud-compression_b200"]
from pcc_b200.train import Trainer
from tools import synth
from tools.bench_ops import timeit
sd = synth.seeded_state_dict(synth.ae_shapes(128, 16, 7), 11)
tr = Trainer(state_dict=sd, ddp=False, device="cuda")
x = torch.from_numpy(synth.modelnet_like(32, 8192, seed=77)).cuda()
start = torch.zeros(32, dtype=torch.int64, device="cuda")
out = {}
b, m = timeit(lambda: out.update(tr.step(x, start)), iters=8, warm=4)
print(f"eager  best {b:.3f} ms  loss {float(out['loss']):.6f}")
b, m = timeit(lambda: out.update(tr.step_graphed(x, start)), iters=10, warm=3)
print(f"graph  best {b:.3f} ms  loss {float(out['loss']):.6f}")
