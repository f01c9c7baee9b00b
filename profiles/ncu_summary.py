"""Summarise an ncu report exported with `--page raw --csv` and `--page source --csv` (run here, no GPU needed):
   python profiles/ncu_summary.py raw.csv src.csv [n_edge_slots]"""
import csv, collections, sys
raw, src = sys.argv[1], sys.argv[2]
nedge = float(sys.argv[3]) if len(sys.argv) > 3 else None
rows=list(csv.reader(open(raw)))
hdr=rows[0]; units=rows[1]; r=rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','l1tex__throughput.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__cycles_elapsed.max','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_fma.sum', 'sm__inst_executed.avg.per_cycle_elapsed','sm__inst_executed_pipe_lsu.sum','l1tex__data_pipe_lsu_wavefronts.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmaheavy.sum','sm__inst_executed_pipe_fmalite.sum']
for i,h in enumerate(hdr):
    if h in want: print(h,units[i],r[i])
for i,h in enumerate(hdr):
    if 'warps_issue_stalled' in h and h.endswith('.ratio'):
        v=float(r[i])
        if v>0.15: print('stall', h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''), r[i])
rows=list(csv.reader(open(src)))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}; data=rows[2:]
tot=sum(int(r[ix['# Samples']]) for r in data)
print('total samples',tot)
for r in sorted(data,key=lambda r:-int(r[ix['# Samples']]))[:22]:
    print(r[ix['# Samples']].rjust(6), r[ix['Instructions Executed']].rjust(9), 'lsb',r[ix['stall_long_sb']].rjust(5),'wait',r[ix['stall_wait']].rjust(5),'ssb',r[ix['stall_short_sb']].rjust(5),'mth',r[ix['stall_math']].rjust(5),'mio',r[ix['stall_mio']].rjust(4), r[ix['Source']].strip()[:80])
ops=collections.Counter()
for r in data:
    toks=r[ix['Source']].strip().split()
    op=toks[0] if not toks[0].startswith('@') else toks[1]
    ops[op.split('.')[0]]+=int(r[ix['Instructions Executed']])
t=sum(ops.values())
if nedge:
    print('instr per warp-edge', t/nedge)
    print(' '.join(f"{k}:{v/nedge:.1f}" for k,v in ops.most_common(28)))
