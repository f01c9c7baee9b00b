import sys, os, time, torch
sys.path.insert(0, os.getcwd())
import numpy as np
import hermnet_b200 as H
from hermnet_b200 import synthetic, ops
name = sys.argv[1] if len(sys.argv) > 1 else "C4"
(pos, Z, cell), cfg = synthetic.config(name, 1.0)
kind = cfg.pop("kind")
torch.manual_seed(1234)
dev = torch.device("cuda:0")
model = getattr(H, kind)(**cfg).to(dev).eval()
for p in model.parameters(): p.requires_grad_(False)
pos_d, Z_d, cell_d = torch.from_numpy(pos).to(dev), torch.from_numpy(Z).to(dev), torch.from_numpy(cell)[None].to(dev)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    g = model.build_graph(pos_d, Z_d, cell_d, None)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("build_graph ms", (t1 - t0) * 1e3, flush=True)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    g = model.build_graph(pos_d, Z_d, cell_d, None)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=15, max_name_column_width=60))
