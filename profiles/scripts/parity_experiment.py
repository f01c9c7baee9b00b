import sys, json, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import cutout, util
from hermnet_b200 import synthetic
import hermnet_b200 as H
(pos, Z, cell), cfg = synthetic.config("C4"); cfg = dict(cfg); kind = cfg.pop("kind")
dev = "cuda"
pos, Z, cell = torch.from_numpy(pos).to(dev), torch.from_numpy(Z).to(dev), torch.from_numpy(cell)[None].to(dev)
for seed_mode in ("torch1234", "oracle1234"):
    if seed_mode == "torch1234":
        torch.manual_seed(1234); model = getattr(H, kind)(**cfg).to(dev).eval()
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    else:
        model, sd = util.make_model(kind, cfg, 1234, dev)
    for p in model.parameters(): p.requires_grad_(False)
    g = model.build_graph(pos, Z, cell)
    for ro in (False, True):
        model.readout_fp32 = ro
        out = cutout.cutout_parity(model, sd, cfg, pos, Z, cell, r_in=4.0, graph=g, check_full_forces=False)
        print(seed_mode, "readout_fp32", ro, "rel_dE", out["rel_dE"], "E", out["E_region"], "dF", out["max_dF"], flush=True)
