"""Summarise an .ncu-rep: key raw metrics + instruction/stall hot spots (needs ncu on PATH)."""
import csv, subprocess, sys, io

KEYS = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__block_size','launch__grid_size',
 'launch__shared_mem_per_block_dynamic','launch__waves_per_multiprocessor','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps',
 'sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
 'smsp__inst_executed.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct',
 'smsp__warps_eligible.avg.per_cycle_active','smsp__warps_active.avg.per_cycle_active','sm__cycles_elapsed.avg',
 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__average_warp_latency_issue_stalled_barrier_per_warp_active.pct']

def raw(rep):
    out = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = [('Kernel Name','',[r[hdr.index('Kernel Name')] for r in rows[2:]])]
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k); res.append((k, units[i], [r[i] for r in rows[2:]]))
    stall = [(h, units[i], [r[i] for r in rows[2:]]) for i,h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
    return res, stall

def source(rep, top=14):
    out = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ai, si, ii, sm = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    stall_cols = [i for i,h in enumerate(hdr) if h.startswith('stall_') and '(Not Issued)' not in h]
    data = []
    for r in rows[2:]:
        if len(r) < len(hdr) or r[ai] in ('Address',''): break
        try: data.append((int(r[ai],16) if r[ai].startswith('0x') else int(r[ai]), r[si], int(r[ii]), int(r[sm]), [int(r[i] or 0) for i in stall_cols]))
        except ValueError: break
    tot = sum(d[2] for d in data); tots = sum(d[3] for d in data); base = data[0][0]
    regions, prev = [], None
    for a,s,c,smp,st in data:
        lvl = c
        if prev is None or abs(lvl-prev['lvl']) > max(2e6, 0.08*prev['lvl']):
            prev = dict(lvl=lvl,a0=a,a1=a,c=c,smp=smp,n=1,st=list(st)); regions.append(prev)
        else:
            prev['a1']=a; prev['c']+=c; prev['smp']+=smp; prev['n']+=1; prev['st']=[x+y for x,y in zip(prev['st'],st)]
    lines = [f"total warp-instructions {tot}, samples {tots}"]
    names = [hdr[i] for i in stall_cols]
    for r in regions:
        if r['c']/tot > 0.004 or r['smp']/tots > 0.01:
            top3 = sorted(zip(r['st'],names), reverse=True)[:3]
            lines.append(f"sass {r['a0']-base:6x}-{r['a1']-base:6x} n={r['n']:4d} exec/instr~{r['lvl']/1e6:8.1f}M instr={r['c']/tot*100:5.1f}% samples={r['smp']/tots*100:5.1f}%  top stalls: " + ", ".join(f"{n}={v/max(r['smp'],1)*100:.0f}%" for v,n in top3))
    allst = [sum(d[4][i] for d in data) for i in range(len(stall_cols))]
    lines.append("stall mix: " + ", ".join(f"{n}={v/tots*100:.1f}%" for v,n in sorted(zip(allst,names), reverse=True)[:8]))
    return lines

if __name__ == '__main__':
    rep = sys.argv[1]
    res, stall = raw(rep)
    for k,u,v in res: print(f"{k},{u},{','.join(v)}")
    for k,u,v in sorted(stall, key=lambda t:-float(t[2][0] or 0))[:8]: print(f"{k},{u},{','.join(v)}")
    if '--no-source' not in sys.argv:
        for l in source(rep): print(l)
