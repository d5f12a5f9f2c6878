"""torch.profiler table of one eager IPDAE train step (32 clouds x 8192): which aten ops / pcc kernels the step's device time goes to."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "point-cloud-compression_b200")]
from pcc_b200.train import Trainer
from tools import synth
from torch.profiler import profile, ProfilerActivity
sd = synth.seeded_state_dict(synth.ae_shapes(128, 16, 7), 11)
tr = Trainer(state_dict=sd, ddp=False, device="cuda")
x = torch.from_numpy(synth.modelnet_like(32, 8192, seed=77)).cuda()
start = torch.zeros(32, dtype=torch.int64, device="cuda")
for _ in range(3):
    tr.step(x, start)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    tr.step(x, start)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
# the aten ops with their input shapes: which tensors the glue kernels (copies, casts, cat) move
print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=30, max_name_column_width=40,
                                                         max_shapes_column_width=90))
