import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
def col(n): return hdr.index(n) if n in hdr else -1
want=[('Kernel Name','name'),('gpu__time_duration.sum','us'),('launch__grid_size','grid'),('launch__block_size','blk'),('launch__registers_per_thread','regs'),
('sm__warps_active.avg.pct_of_peak_sustained_active','occ%'),('dram__bytes_read.sum','rdMB'),('dram__bytes_write.sum','wrMB'),('dram__throughput.avg.pct_of_peak_sustained_elapsed','dram%'),
('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','fma%'),('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','tensor%'),('smsp__issue_active.avg.pct_of_peak_sustained_active','issue%'),('lts__t_sector_hit_rate.pct','l2hit%'),
('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','st_long'),('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','st_wait'),('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','st_bar'),('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','st_lg'),('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','st_mio'),('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','st_short'),('smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','st_noinst'),('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','st_math')]
units=rows[1]
for r in rows[2:]:
    parts=[]
    for n,a in want:
        i=col(n)
        if i<0: continue
        v=r[i]
        if a=='name': v=v[:44]
        elif a in('rdMB','wrMB'):
            u=units[i]; f=float(v.replace(',',''))
            f = f/1e6 if u=='byte' else (f/1e3 if u=='Kbyte' else (f if u=='Mbyte' else f*1e3))
            v=f'{f:.0f}'
        elif a=='us':
            u=units[i]; f=float(v.replace(',','')); f = f/1e3 if u=='ns' else (f if u=='us' else f*1e3); v=f'{f:.1f}'
        else:
            try: v=f'{float(v.replace(",","")):.2f}'
            except: pass
        parts.append(f'{a}={v}')
    print(' '.join(parts))
