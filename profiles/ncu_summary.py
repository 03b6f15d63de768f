import csv, sys, subprocess, re
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
h=rows[0]
def col(n): return h.index(n) if n in h else None
M=[('gpu__time_duration.sum','us'),('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','tensor%'),
   ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','dram%'),('dram__bytes_read.sum','rdMB'),('dram__bytes_write.sum','wrMB'),
   ('lts__throughput.avg.pct_of_peak_sustained_elapsed','l2%'),('l1tex__throughput.avg.pct_of_peak_sustained_elapsed','l1%'),
   ('sm__throughput.avg.pct_of_peak_sustained_elapsed','sm%'),('sm__warps_active.avg.pct_of_peak_sustained_active','occ%'),
   ('launch__registers_per_thread','regs'),('smsp__inst_executed.sum','inst'),('launch__grid_size','grid'),
   ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','st_long'),
   ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','st_short'),
   ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','st_bar'),
   ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','st_mio'),
   ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','st_math'),
   ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','st_wait'),
   ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','bankconf')]
seen={}
for r in rows[2:]:
    name=re.sub(r'\(.*','',r[col('Kernel Name')]); name=re.sub(r'.*::','',name)
    key=name
    seen[key]=seen.get(key,0)+1
    if seen[key]>int(sys.argv[2]) if len(sys.argv)>2 else 2: continue
    vals=[]
    for m,lab in M:
        c=col(m)
        if c is None: continue
        v=r[c]
        try: v=float(v); v=f"{v:.3g}"
        except: pass
        vals.append(f"{lab}={v}")
    print(name[:40].ljust(40),' '.join(vals))
