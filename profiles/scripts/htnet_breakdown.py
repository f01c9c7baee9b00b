"""Where an HTNet evaluation at C3 size goes: per-kernel CUDA-event times of our launches + total step time."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import hermnet_b200 as H
from hermnet_b200 import ops, synthetic
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
(pos, Z, cell), cfg = synthetic.config(name, float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
kind = cfg.pop("kind")
dev = "cuda"
torch.manual_seed(0)
model = getattr(H, kind)(**cfg).to(dev).eval()
for p in model.parameters(): p.requires_grad_(False)
pos, Z, cell = torch.from_numpy(pos).to(dev), torch.from_numpy(Z).to(dev), torch.from_numpy(cell)[None].to(dev)
g = model.build_graph(pos, Z, cell)
def step():
    p = pos.detach().requires_grad_(True)
    e, _, _ = model.forward_graph(p, Z, cell, g)
    (gr,) = torch.autograd.grad(e.sum(), p)
for _ in range(3): step()
torch.cuda.synchronize()
ops.TIMERS = {}; ops.LAUNCHES["n"] = 0
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): step()
b.record(); torch.cuda.synchronize()
tm, ops.TIMERS = ops.TIMERS, None
ms = a.elapsed_time(b) / 5
rows = {k: (len(v) // 5, sum(x.elapsed_time(y) for x, y in v) / 5) for k, v in tm.items()}
print(f"{name} {kind} N={len(Z)} rows={g.n_rows} row-edges={g.n_edges} modules={g.n_modules}: {ms:.2f} ms/step, our launches/step {ops.LAUNCHES['n'] // 5}, "
      f"sum of our kernels {sum(v[1] for v in rows.values()):.2f} ms")
for nm in ("tc_dst", "tc_src"):
    pl = g._lazy.get(nm)
    if pl is not None:
        cnt = pl.tile_info[: pl.n_tiles, 1].float()
        print(f"   plan {nm}: window {pl.window}, blocks {pl.n_blocks}, tiles {pl.n_tiles}, fill mean {float(cnt.mean()):.1f}, chunks mean "
              f"{float(pl.tile_win[: pl.n_tiles, 1].float().mean()):.2f}")
for k, v in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    print(f"   {k:24s} {v[0]:5d} launches  {v[1]:8.2f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
