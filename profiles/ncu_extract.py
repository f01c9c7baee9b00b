"""Summarise an `ncu --set full` report of the edge kernels (run here, no GPU needed):

    python profiles/ncu_extract.py gpurun_out/prof.ncu-rep profiles/ncu_r1_vN [n_edges]

writes  <out>_summary.md   one row per captured launch: time, DRAM bytes, L1/L2 hit rates, pipe utilisation, IPC,
                           registers, instructions per edge, dominant stall reasons
        <out>_raw.csv      the metrics the table was built from (one line per launch)
        profiles/traffic.json   bench.py's `roofline.traffic`: mean dram__bytes_read.sum + dram__bytes_write.sum per
                           launch of every edge kernel (keyed by the timer names of hermnet_b200/ops.py)
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "GB rd"), ("dram__bytes_write.sum", "GB wr"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->L1 GB"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data pipe %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp instr"),
]
TIMER = {"edge_fwd": "painn_edge_fwd", "edge_bwd_dst": "painn_edge_bwd_dst", "edge_bwd_src": "painn_edge_bwd_src",
         "edge_tc_fwd": "painn_edge_fwd", "edge_tc_bwd_dst": "painn_edge_bwd_dst", "edge_tc_bwd_src": "painn_edge_bwd_src"}


def to_gb(v, unit):
    v = float(v)
    return v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}.get(unit, 1.0)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    n_edges = float(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    table, rawrows, traffic = [], [], collections.defaultdict(list)
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        short = name.split("::")[-1].split("(")[0]
        m = re.search(r"edge_tc_kernel<(?:\(int\))?(\d), (?:\(bool\))?(\d|true|false)", name)
        if m:       # tensor-core edge kernels: template <MODE, HV, DBG>
            short = ["edge_tc_fwd", "edge_tc_bwd_dst", "edge_tc_bwd_src"][int(m.group(1))] + ("" if m.group(2) in ("1", "true") else "[vec=NULL]")
        rec = {"kernel": short}
        for key, label in WANT:
            if key not in ix:
                continue
            u, v = units[ix[key]], r[ix[key]]
            rec[label] = to_gb(v, u) if "byte" in u else float(v)
            if key == "gpu__time_duration.sum":
                rec[label] = float(v) * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(u, 1.0)
        stalls = []
        for h, i in ix.items():
            if "warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h.split("stalled_")[1].replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        rec["top stalls"] = ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:4] if n != "selected")
        if n_edges:
            rec["instr/edge"] = rec.get("warp instr", 0.0) / n_edges
        table.append(rec)
        rawrows.append([short] + [r[ix[k]] for k, _ in WANT if k in ix])
        for frag, timer in TIMER.items():
            if frag in short and "[vec=NULL]" not in short:      # (the first-layer variants read an L1-resident table)
                traffic[timer].append(1e9 * (rec.get("GB rd", 0.0) + rec.get("GB wr", 0.0)))
    cols = ["kernel"] + [l for _, l in WANT if l != "warp instr"] + (["instr/edge"] if n_edges else []) + ["top stalls"]
    with open(out + "_summary.md", "w") as f:
        f.write(f"ncu --set full, {os.path.basename(rep)} ({len(table)} launches)\n\n")
        f.write("| " + " | ".join(cols) + " |\n|" + "---|" * len(cols) + "\n")
        for rec in table:
            f.write("| " + " | ".join(f"{rec.get(c, ''):.3g}" if isinstance(rec.get(c), float) else str(rec.get(c, "")) for c in cols) + " |\n")
    with open(out + "_raw.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [k for k, _ in WANT if k in ix])
        w.writerows(rawrows)
    if not traffic:        # no edge kernel in this report: leave bench.py's traffic table alone
        print(open(out + "_summary.md").read())
        return
    tj = {k: sum(v) / len(v) for k, v in traffic.items()}
    tj["_source"] = f"{os.path.basename(rep)}: mean dram__bytes_read.sum + dram__bytes_write.sum per launch (bytes)"
    json.dump(tj, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json"), "w"), indent=1)
    print(open(out + "_summary.md").read())


if __name__ == "__main__":
    main()
