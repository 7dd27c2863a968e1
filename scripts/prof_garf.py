import os, sys
sys.path.insert(0, "/root/repo")
import torch as th
from nerf_experiments_b200.model_garf import GarfModel
dev = th.device("cuda:0")
th.manual_seed(1337)
m = GarfModel(2.0, 7.0, 64, 192, 0.5, 1.5, 1.0, 1e-3, 1e-4, 100000, 0.0, 1e-3, 1e-4, 100000, 0.0).to(dev)
m.train()
B = 1024
g = th.Generator().manual_seed(0)
o = (th.nn.functional.normalize(th.randn((B, 3), generator=g), dim=1) * 4.0).to(dev)
d = th.nn.functional.normalize(-o.cpu() + 0.3 * th.randn((B, 3), generator=g), dim=1).to(dev)
tgt = th.rand((B, 3), generator=g).to(dev)
for i in range(3):
    m.training_step((o, d, tgt), i)
th.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3):
        m.training_step((o, d, tgt), 3 + i)
    th.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
