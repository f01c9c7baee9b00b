"""Summarise an ``ncu --metrics gpu__time_duration.sum --csv`` launch list: per-kernel totals and shares.
Usage: python profiles/analyze_launches.py gpurun_out/launches.csv [skip_fraction]"""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    out = []
    for x in rows:
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        v = v / 1e6 if u == "ns" else (v / 1e3 if u.startswith("us") else v)
        name = x["Kernel Name"]
        t = re.search(r"edge_tc_kernel<(?:\(int\))?(\d), (?:\(bool\))?(\d|true|false)", name)
        if t:
            out.append((["edge_tc_fwd", "edge_tc_bwd_dst", "edge_tc_bwd_src"][int(t.group(1))], v))
            continue
        t = re.search(r"(layer0_basis_kernel|layernorm_kernel|readout_kernel)<(?:\(bool\))?(?:\d+, )?(?:\(bool\))?(\d|true|false)>", name)
        if t:
            out.append((t.group(1).replace("_kernel", "") + ("_bwd" if t.group(2) in ("1", "true") else "_fwd"), v))
            continue
        m = re.search(r"(gemm_tf32x3_kernel|node_\w+_kernel|halo_\w+_kernel|tile_windows_kernel|split_weights_kernel|plan_\w+_kernel|"
                      r"basis_index_kernel|split_tf32\w*_kernel)", name)
        if m:
            out.append((m.group(1), v))
            continue
        m = re.search(r"(edge_\w+_kernel|rg_\w+_kernel|segment_sum\w*_kernel|gather_rows\w*_kernel|expand_rowptr_kernel|"
                      r"key_hist_kernel|graph_\w+_kernel|bin_atoms_kernel|gather_sorted_kernel|triplet\w+_kernel)", name)
        short = m.group(1) if m else re.sub(r"<.*", "", name).replace("void ", "")[:60]
        out.append((short, v))
    return out


def main():
    rows = load(sys.argv[1])
    start = 0
    if len(sys.argv) > 2:   # only launches after the k-th edge_geom_fwd_kernel (skip set-up + warm-up steps)
        k, seen = int(sys.argv[2]), 0
        for i, (n, _) in enumerate(rows):
            if n == "edge_geom_fwd_kernel":
                seen += 1
                if seen == k + 1:
                    start = i
                    break
    rows = rows[start:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v in rows:
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"launches {len(rows)}  total {tot:.1f} ms")
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
        print(f"| {k} | {v[0]} | {v[1]:.2f} | {100 * v[1] / tot:.1f}% |")


if __name__ == "__main__":
    main()
