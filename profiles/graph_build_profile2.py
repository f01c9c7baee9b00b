"""Phases of the per-step graph work at C4 size: neighbour list + row CSR, tile plans (destination- and source-major)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hermnet_b200 as H
from hermnet_b200 import synthetic, tileplan, ops, functional as Fn
(pos, Z, cell), cfg = synthetic.config("C4"); kind = cfg.pop("kind")
dev = "cuda"
model = getattr(H, kind)(**cfg).to(dev).eval()
p, z, c = torch.from_numpy(pos).to(dev), torch.from_numpy(Z).to(dev), torch.from_numpy(cell)[None].to(dev)
def tick():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(3):
    t0 = tick(); g = model.build_graph(p, z, c)
    t1 = tick(); geom = Fn.edge_geometry(p[g.perm], c, g)
    t2 = tick(); dst = tileplan.build_dst_plan(g, geom, 0.2, 128)
    t3 = tick(); src = tileplan.build_src_plan(g, geom, 0.2, 128)
    t4 = tick(); dst.update_windows(geom, 0.2, 128); src.update_windows(geom, 0.2, 128)
    t5 = tick(); uniq, g0 = model._layer0_tables(g, z[g.perm].long()); d0 = tileplan.dst_plan_for_table(dst, g, g0)
    t6 = tick()
    print(f"build_graph {1e3*(t1-t0):.1f}  geom {1e3*(t2-t1):.1f}  dst plan {1e3*(t3-t2):.1f}  src plan {1e3*(t4-t3):.1f}  windows x2 {1e3*(t5-t4):.1f}  "
          f"layer-0 tables + plan {1e3*(t6-t5):.1f} ms")
ops.TIMERS = {}
g = model.build_graph(p, z, c); torch.cuda.synchronize()
tm, ops.TIMERS = ops.TIMERS, None
print({k: round(sum(a.elapsed_time(b) for a, b in v), 2) for k, v in tm.items()})
