"""Wall-clock phases of one end-to-end evaluation (H2D, graph build, forward, backward, D2H) for a named workload:
    python profiles/e2e_phases.py C3"""
import os, sys, time
import torch
sys.path.insert(0, os.getcwd())
import hermnet_b200 as H
from hermnet_b200 import synthetic

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
(pos, Z, cell), cfg = synthetic.config(name, float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
kind = cfg.pop("kind")
torch.manual_seed(1234)
dev = torch.device("cuda:0")
model = getattr(H, kind)(**cfg).to(dev).eval()
for p in model.parameters():
    p.requires_grad_(False)
pos_h, Z_h, cell_h = torch.from_numpy(pos).pin_memory(), torch.from_numpy(Z).pin_memory(), torch.from_numpy(cell)[None].pin_memory()


def tick():
    torch.cuda.synchronize()
    return time.perf_counter()


import gc
n_it = int(sys.argv[3]) if len(sys.argv) > 3 else 4
for it in range(n_it):
    t0 = tick()
    p_, z_, c_ = pos_h.to(dev, non_blocking=True), Z_h.to(dev, non_blocking=True), cell_h.to(dev, non_blocking=True)
    t1 = tick()
    g = model.build_graph(p_, z_, c_, None)
    t2 = tick()
    pr = p_.detach().requires_grad_(True)
    e, _, _ = model.forward_graph(pr, z_, c_, g)
    t3 = tick()
    (gr,) = torch.autograd.grad(e.sum(), pr)
    t4 = tick()
    f = (-gr).cpu()
    t5 = tick()
    d = H.Data(pos=p_.detach().requires_grad_(True), atomic_number=z_, cell=c_)
    e2 = model(d)
    t6 = tick()
    (g2,) = torch.autograd.grad(e2.sum(), d.pos)
    t7 = tick()
    print(f"{name} it{it}: h2d {1e3*(t1-t0):.2f}  build_graph {1e3*(t2-t1):.2f}  forward_graph {1e3*(t3-t2):.2f}  backward {1e3*(t4-t3):.2f}  "
          f"d2h {1e3*(t5-t4):.2f} | model(data) {1e3*(t6-t5):.2f}  backward {1e3*(t7-t6):.2f} ms | alloc {torch.cuda.memory_allocated()/2**30:.1f} GiB "
          f"reserved {torch.cuda.memory_reserved()/2**30:.1f} GiB gc {gc.get_count()}", flush=True)
