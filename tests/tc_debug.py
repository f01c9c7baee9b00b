"""Bring-up diagnostics of the tensor-core edge kernels (not a pytest file): per-stage error report against a dense
float64 evaluation and the row-per-warp kernels.  Usage: python tests/tc_debug.py [n_side] [stage ...]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from hermnet_b200 import ops, tileplan  # noqa: E402
from tests.test_sweep_kernels import _dense_reference  # noqa: E402
from tests.util import edge_inputs, frozen_model, lattice_system  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


def main():
    n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    stages = sys.argv[2:] or ["plan", "phi", "fwd", "fwd0", "dst", "dst0", "src"]
    dev = "cuda:0"
    pos, Z, cell = lattice_system(n_side, [3, 13, 14, 8], 5)
    pos, Z, cell = pos.to(dev), Z.to(dev), cell.to(dev)
    model = frozen_model("HVNet", ["Li", "Al", "Si", "O"], 128, 128, dev)
    g = model.build_graph(pos, Z, cell)
    p, geom, xh, vec, Wt, bias, off, g_dx, g_dvec = edge_inputs(model, g, pos, cell)
    print(f"atoms {g.n_atoms} edges {g.n_edges} rows {g.n_rows}", flush=True)
    t0 = time.time()
    dst, src = tileplan.plans_of(g, geom, p.inv_rc, p.num_rbf, want_src=True)
    torch.cuda.synchronize()
    print(f"plans built in {time.time() - t0:.3f}s: dst blocks {dst.n_blocks} tiles {dst.n_tiles}; src blocks {src.n_blocks} "
          f"tiles {src.n_tiles}", flush=True)
    dst.update_windows(geom, p.inv_rc, p.num_rbf)
    src.update_windows(geom, p.inv_rc, p.num_rbf)
    torch.cuda.synchronize()
    if "plan" in stages:
        for name, pl in (("dst", dst), ("src", src)):
            ti = pl.tile_info[: pl.n_tiles].cpu()
            win = pl.tile_win[: pl.n_tiles].cpu()
            er = pl.erec.cpu()
            n_live = int((g.row_mod.long()[g.edge_row.long()] >= 0).sum())
            cnt = ti[:, 1]
            print(f"  {name}: tile fill mean {float(cnt.float().mean()):.1f} max {int(cnt.max())} min {int(cnt.min())}; "
                  f"chunks max {int(win[:, 1].max())} mean {float(win[:, 1].float().mean()):.3f}; "
                  f"edges in tiles {int(cnt.sum())} live {n_live}; unique eids {int(torch.unique(er[: int(cnt.sum()), 3]).numel())}",
                  flush=True)
    big = n_side > 8
    if big:      # timing / plan statistics only: compare with the tile-sweep kernels instead of float64
        wsplit, wscale = ops.tc_split_weights(Wt)
        for name, a, b in (("fwd", lambda: ops.tc_edge_fwd(p, dst, xh, vec, geom, wsplit, wscale, bias, off, p.n_rows),
                            lambda: ops.painn_edge_fwd(p, xh, vec, geom, g, Wt, bias, off)),
                           ("dst", lambda: ops.tc_edge_bwd_dst(p, dst, xh, vec, geom, wsplit, wscale, bias, off, g_dx, g_dvec)[0],
                            lambda: ops.painn_edge_bwd_dst(p, xh, vec, geom, g, Wt, bias, off, g_dx, g_dvec).sum(0)),
                           ("src", lambda: ops.tc_edge_bwd_src(p, src, xh, vec, geom, wsplit, wscale, bias, off, g_dx, g_dvec),
                            lambda: ops.painn_edge_bwd_src(p, xh, vec, geom, g, Wt, bias, off, g_dx, g_dvec))):
            if name not in stages:
                continue
            ra, rb = a(), b()
            torch.cuda.synchronize()
            ra, rb = (ra if isinstance(ra, tuple) else (ra,)), (rb if isinstance(rb, tuple) else (rb,))
            print(name, "tc vs quad:", [f"{rel(x, y.view_as(x)):.3e}" for x, y in zip(ra, rb)], flush=True)

        def timeit(fn, n=5):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / n
        res = {}
        if "fwd" in stages:
            res["tc fwd"] = timeit(lambda: ops.tc_edge_fwd(p, dst, xh, vec, geom, wsplit, wscale, bias, off, p.n_rows))
            res["tc fwd0"] = timeit(lambda: ops.tc_edge_fwd(p, dst, xh, None, geom, wsplit, wscale, bias, off, p.n_rows))
            res["quad fwd"] = timeit(lambda: ops.painn_edge_fwd(p, xh, vec, geom, g, Wt, bias, off))
        if "dst" in stages:
            res["tc bwd_dst"] = timeit(lambda: ops.tc_edge_bwd_dst(p, dst, xh, vec, geom, wsplit, wscale, bias, off, g_dx, g_dvec))
            res["quad bwd_dst"] = timeit(lambda: ops.painn_edge_bwd_dst(p, xh, vec, geom, g, Wt, bias, off, g_dx, g_dvec))
        if "src" in stages:
            res["tc bwd_src"] = timeit(lambda: ops.tc_edge_bwd_src(p, src, xh, vec, geom, wsplit, wscale, bias, off, g_dx, g_dvec))
            res["quad bwd_src"] = timeit(lambda: ops.painn_edge_bwd_src(p, xh, vec, geom, g, Wt, bias, off, g_dx, g_dvec))
        print("ms per launch:", {k: round(v, 3) for k, v in res.items()}, "edges", g.n_edges, flush=True)
        return
    ref = _dense_reference(g, p, geom, xh, vec, Wt, bias, off, g_dx, g_dvec)
    ops.edge_set_variant("row")
    row_dx, row_dv = ops.painn_edge_fwd(p, xh, vec, geom, g, Wt, bias, off)
    row_gg = ops.painn_edge_bwd_dst(p, xh, vec, geom, g, Wt, bias, off, g_dx, g_dvec).sum(0)
    row_gxh, row_gvec = ops.painn_edge_bwd_src(p, xh, vec, geom, g, Wt, bias, off, g_dx, g_dvec)
    row_dx0, row_dv0 = ops.painn_edge_fwd(p, xh, None, geom, g, Wt, bias, off)
    row_gg0 = ops.painn_edge_bwd_dst(p, xh, None, geom, g, Wt, bias, off, g_dx, g_dvec).sum(0)
    ops.edge_set_variant("auto")
    torch.cuda.synchronize()
    wsplit, wscale = ops.tc_split_weights(Wt)
    torch.cuda.synchronize()
    print("weights split; wscale", wscale.tolist(), flush=True)
    if "phi" in stages or "fwd" in stages:
        dx, dv, dbg = ops.tc_edge_fwd(p, dst, xh, vec, geom, wsplit, wscale, bias, off, p.n_rows, debug_phi=True)
        torch.cuda.synchronize()
        E3 = g.n_edges * 384
        phi, raw = dbg[:E3].view(-1, 384), dbg[E3:2 * E3].view(-1, 384)
        dump = dbg[2 * E3:].view(torch.float16).view(-1, 32).cpu()          # rows of 64 bytes (64-byte swizzle)
        print(f"nan count phi {int(torch.isnan(phi).sum())} of {phi.numel()}, raw {int(torch.isnan(raw).sum())}, inf raw "
              f"{int(torch.isinf(raw).sum())}; dx nan {int(torch.isnan(dx).sum())}", flush=True)

        def unswizzle(rows):
            r = torch.arange(rows.size(0))
            out = torch.empty_like(rows)
            v = rows.view(-1, 4, 8)
            for c in range(4):
                out.view(-1, 4, 8)[:, c] = v[r, c ^ ((r >> 1) & 3)]
            return out
        tiles = unswizzle(dump[:896])
        # W: tile (part, half) = rows (2 part + half) * 128 ..; put back as [384 channels, hi 32 | lo 32]
        wd = torch.cat([torch.cat([tiles[(2 * pp_) * 128:(2 * pp_ + 1) * 128], tiles[(2 * pp_ + 1) * 128:(2 * pp_ + 2) * 128]], 1)
                        for pp_ in range(3)], 0)
        ed = torch.cat([tiles[768:832], tiles[832:896]], 1)                  # [64 edges, hi 32 | lo 32]
        ti0 = dst.tile_info[int(dst.blk_tile[0])].tolist()
        k0 = int(dst.tile_win[int(dst.blk_tile[0]), 0])
        m0 = ti0[3]
        K32 = 128
        ws = wsplit.view(2, -1, K32).cpu().permute(1, 0, 2)                   # [rows, 2, K32]
        w_exp = torch.cat([ws[m0 * 384:(m0 + 1) * 384, 0, k0:k0 + 32], ws[m0 * 384:(m0 + 1) * 384, 1, k0:k0 + 32]], 1)
        print(f"first tile: start {ti0[0]} cnt {ti0[1]} even {ti0[2]} mod {m0} k0 {k0}; W stage vs expected max diff "
              f"{float((wd.float() - w_exp.float()).abs().max()):.3e} (nan in dump {int(torch.isnan(wd.float()).sum())})", flush=True)
        print(f"  wsplit: nan {int(torch.isnan(ws.float()).sum())} max |hi| {float(ws[:, 0].float().abs().max()):.1f} max |lo| "
              f"{float(ws[:, 1].float().abs().max()):.3f}", flush=True)
        for part in range(3):
            for half in range(2):
                a = wd[part * 128:(part + 1) * 128, half * 32:(half + 1) * 32].float()
                b = w_exp[part * 128:(part + 1) * 128, half * 32:(half + 1) * 32].float()
                bad = torch.nonzero(~(a == b))
                print(f"  W part {part} half {half}: mismatches {bad.size(0)} of {a.numel()}"
                      + (f"; first at (row {int(bad[0, 0])}, k {int(bad[0, 1])}): got {float(a[bad[0, 0], bad[0, 1]])} exp "
                         f"{float(b[bad[0, 0], bad[0, 1]])}; rows hit {sorted(set(bad[:, 0].tolist()))[:12]} cols "
                         f"{sorted(set(bad[:, 1].tolist()))[:12]}" if bad.size(0) else ""), flush=True)
        er = dst.erec[ti0[0]:ti0[0] + ti0[1]].cpu().long()
        ue = geom[er[:, 3].to(geom.device), 3].cpu().double() * float(p.inv_rc)
        pe_ = p.env_p
        a_, b_, c_ = -(pe_ + 1) * (pe_ + 2) / 2, pe_ * (pe_ + 2), -pe_ * (pe_ + 1) / 2
        env_ = torch.where(ue < 1, 1 + a_ * ue ** pe_ + b_ * ue ** (pe_ + 1) + c_ * ue ** (pe_ + 2), torch.zeros_like(ue))
        kk = torch.arange(k0, k0 + 32)
        offc = off.cpu().double()
        val = env_[:, None] * torch.exp(float(p.coeff) * (ue[:, None] - offc[kk.clamp(max=127)][None, :]) ** 2) * 2048.0
        val = torch.where(kk[None, :] < 128, val, torch.zeros_like(val))
        got = ed[: ti0[1], :32].double() + ed[: ti0[1], 32:].double()
        print(f"  basis operand vs expected: max abs diff {float((got - val).abs().max()):.3e} (scale 2048); rows beyond cnt max "
              f"{float(ed[ti0[1]:].float().abs().max()) if ti0[1] < 64 else 0.0:.3e}", flush=True)
        dd_ = (got - val).abs()
        ij = torch.nonzero(dd_ > 1e-2)
        print(f"  basis mismatches {ij.size(0)} of {dd_.numel()}" + (f"; first (row {int(ij[0, 0])}, k {int(ij[0, 1])}) got "
              f"{float(got[ij[0, 0], ij[0, 1]])} exp {float(val[ij[0, 0], ij[0, 1]])}; rows {sorted(set(ij[:, 0].tolist()))[:16]} cols "
              f"{sorted(set(ij[:, 1].tolist()))[:16]}" if ij.size(0) else ""), flush=True)
        # raw accumulators of the tile's first edge vs the product of the dumped operands
        prod = (wd[:, :32].double() @ ed[:, :32].double().t() + wd[:, 32:].double() @ ed[:, :32].double().t()
                + wd[:, :32].double() @ ed[:, 32:].double().t())           # [384, 64]
        e0 = int(er[0, 3])
        print("  raw acc of first record (ch 0..5):", raw[e0, :6].tolist(), "expected", prod[:6, 0].tolist(), flush=True)
        print("  raw acc part b/c (ch 128.., 256..):", raw[e0, 128:131].tolist(), prod[128:131, 0].tolist(), raw[e0, 256:259].tolist(),
              prod[256:259, 0].tolist(), flush=True)
        rawt = raw[er[:, 3].to(raw.device)].cpu().double().t()                 # [384, cnt]
        print(f"  raw acc of the whole first tile vs operand product: max abs diff {float((rawt - prod[:, : ti0[1]]).abs().max()):.3e} "
              f"(scale {float(prod.abs().max()):.3e})", flush=True)
        # dense float64 phi
        row = g.edge_row.long()
        m = g.row_mod.long()[row].clamp(min=0)
        u = geom[:, 3].double() * float(p.inv_rc)
        pe = p.env_p
        a, b, c = -(pe + 1) * (pe + 2) / 2, pe * (pe + 2), -pe * (pe + 1) / 2
        env = torch.where(u < 1, 1 + a * u ** pe + b * u ** (pe + 1) + c * u ** (pe + 2), torch.zeros_like(u))
        emb = env[:, None] * torch.exp(float(p.coeff) * (u[:, None] - off.double()[None, :]) ** 2)
        phi_ref = torch.einsum("ek,ekc->ec", emb, Wt.double()[m]) + bias.double()[m]
        live = (g.row_mod.long()[row] >= 0)
        err = (phi.double() - phi_ref).abs()[live]
        print(f"phi: max abs err {float(err.max()):.3e} (scale {float(phi_ref.abs().max()):.3e}); per part "
              f"{[float(err[:, i * 128:(i + 1) * 128].max()) for i in range(3)]}", flush=True)
        if float(err.max()) > 1e-3:
            bad = torch.nonzero(err.max(1).values > 1e-3).squeeze(1)
            print(f"  bad edges {bad.numel()} of {err.size(0)}; first {bad[:10].tolist()}", flush=True)
            e = int(torch.nonzero(live).squeeze(1)[bad[0]])
            print("  edge", e, "u", float(u[e]), "phi", phi[e, :6].tolist(), "ref", phi_ref[e, :6].tolist(), flush=True)
        print(f"fwd: dx rel {rel(dx, ref[0]):.3e} (row kernels {rel(row_dx, ref[0]):.3e}); dvec rel {rel(dv, ref[1]):.3e} "
              f"(row {rel(row_dv, ref[1]):.3e})", flush=True)
    if "fwd0" in stages:
        dx, dv = ops.tc_edge_fwd(p, dst, xh, None, geom, wsplit, wscale, bias, off, p.n_rows)
        torch.cuda.synchronize()
        print(f"fwd vec=NULL: dx vs row {rel(dx, row_dx0):.3e}; dvec vs row {rel(dv, row_dv0):.3e}", flush=True)
    if "dst" in stages:
        gg = ops.tc_edge_bwd_dst(p, dst, xh, vec, geom, wsplit, wscale, bias, off, g_dx, g_dvec)[0]
        torch.cuda.synchronize()
        print(f"bwd_dst: g_geom rel {rel(gg, ref[2]):.3e} (row {rel(row_gg, ref[2]):.3e}); d-part {rel(gg[:, 3], ref[2][:, 3]):.3e} "
              f"u-part {rel(gg[:, :3], ref[2][:, :3]):.3e}", flush=True)
    if "dst0" in stages:
        gg = ops.tc_edge_bwd_dst(p, dst, xh, None, geom, wsplit, wscale, bias, off, g_dx, g_dvec)[0]
        torch.cuda.synchronize()
        print(f"bwd_dst vec=NULL: vs row {rel(gg, row_gg0):.3e}", flush=True)
    if "src" in stages:
        gxh, gvec = ops.tc_edge_bwd_src(p, src, xh, vec, geom, wsplit, wscale, bias, off, g_dx, g_dvec)
        torch.cuda.synchronize()
        print(f"bwd_src: grad_xh rel {rel(gxh, ref[3]):.3e} (row {rel(row_gxh, ref[3]):.3e}); grad_vec rel "
              f"{rel(gvec, ref[4].view_as(gvec)):.3e} (row {rel(row_gvec, ref[4].view_as(row_gvec)):.3e})", flush=True)
    if "time" in stages:
        def timeit(fn, n=5):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / n
        print("ms: tc fwd", timeit(lambda: ops.tc_edge_fwd(p, dst, xh, vec, geom, wsplit, wscale, bias, off, p.n_rows)),
              "tc bwd_dst", timeit(lambda: ops.tc_edge_bwd_dst(p, dst, xh, vec, geom, wsplit, wscale, bias, off, g_dx, g_dvec)),
              "tc bwd_src", timeit(lambda: ops.tc_edge_bwd_src(p, src, xh, vec, geom, wsplit, wscale, bias, off, g_dx, g_dvec)),
              "quad fwd", timeit(lambda: ops.painn_edge_fwd(p, xh, vec, geom, g, Wt, bias, off)),
              "quad bwd_dst", timeit(lambda: ops.painn_edge_bwd_dst(p, xh, vec, geom, g, Wt, bias, off, g_dx, g_dvec)),
              "quad bwd_src", timeit(lambda: ops.painn_edge_bwd_src(p, xh, vec, geom, g, Wt, bias, off, g_dx, g_dvec)), flush=True)


if __name__ == "__main__":
    main()
