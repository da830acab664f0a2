import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
keys=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','lts__t_sectors_op_read.sum','lts__t_sectors_op_red.sum','lts__t_sectors_op_atom.sum','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit_registers','smsp__inst_executed.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__data_pipe_lsu_wavefronts.sum','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed','l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__thread_inst_executed_per_inst_executed.ratio','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__cycles_active.avg','sm__cycles_elapsed.max']
stall=[h for h in hdr if "average_warps_issue_stalled" in h]
for r in rows[2:]:
    for k in keys:
        if k in hdr: print(k,'=',r[hdr.index(k)],rows[1][hdr.index(k)])
    st=sorted(((float(r[hdr.index(k)] or 0),k) for k in stall),reverse=True)[:7]
    for v,k in st: print('   stall',k.replace("smsp__average_warps_issue_stalled_","").replace("_per_issue_active.ratio",""),round(v,1))
    print('---')
