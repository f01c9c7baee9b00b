import sys, time, torch
sys.path.insert(0, ".")
import hermnet_b200 as H
from hermnet_b200 import synthetic, ops
(pos, Z, cell), cfg = synthetic.config("C4"); kind = cfg.pop("kind")
dev = torch.device("cuda:0")
torch.manual_seed(1234)
model = getattr(H, kind)(**cfg).to(dev).eval()
for p in model.parameters(): p.requires_grad_(False)
pos_h, Z_h, cell_h = torch.from_numpy(pos).pin_memory(), torch.from_numpy(Z).pin_memory(), torch.from_numpy(cell)[None].pin_memory()
f_h = torch.empty((len(Z), 3)).pin_memory(); e_h = torch.empty(1).pin_memory()
resident = "--resident" in sys.argv
if resident:
    pd, zd, cd = pos_h.to(dev), Z_h.to(dev), cell_h.to(dev)
    graph = model.build_graph(pd, zd, cd, None)
    for _ in range(4):
        p = pd.detach().requires_grad_(True); e, _, _ = model.forward_graph(p, zd, cd, graph); torch.autograd.grad(e.sum(), p)
for it in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    p = pos_h.to(dev, non_blocking=True); z = Z_h.to(dev, non_blocking=True); c = cell_h.to(dev, non_blocking=True)
    d = H.Data(pos=p.requires_grad_(True), atomic_number=z, cell=c)
    e = model(d)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    (g,) = torch.autograd.grad(e.sum(), d.pos)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    f_h.copy_(-g, non_blocking=True); e_h.copy_(e.detach().reshape(1), non_blocking=True)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"it{it}: fwd {1e3*(t1-t0):.1f} bwd {1e3*(t2-t1):.1f} d2h {1e3*(t3-t2):.1f} total {1e3*(t3-t0):.1f} alloc {torch.cuda.memory_allocated()/2**30:.1f} max {torch.cuda.max_memory_allocated()/2**30:.1f} reserved {torch.cuda.memory_reserved()/2**30:.1f} GiB", flush=True)
    torch.cuda.reset_peak_memory_stats()
    if "--del" in sys.argv:
        del d, e, g, p, z, c
