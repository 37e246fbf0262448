import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
r=list(csv.reader(out.splitlines()))
h=r[0]
want=['gpu__time_duration.sum','smsp__inst_executed.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__registers_per_thread','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__cycles_active.avg','launch__waves_per_multiprocessor','dram__bytes_read.sum','dram__bytes_write.sum']
for v in r[2:]:
    print(v[h.index('Kernel Name')][:60])
    for w in want:
        if w in h: print('  ',w, v[h.index(w)])
    for i,n in enumerate(h):
        if 'issue_stalled' in n and n.endswith('per_issue_active.ratio'):
            try:
                if float(v[i])>0.15: print('   stall',n.split('issue_stalled_')[1].split('_per_')[0], v[i])
            except: pass
