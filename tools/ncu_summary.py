import sys, csv, subprocess
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts.sum','smsp__inst_executed.sum','sm__warps_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','dram__throughput.avg.pct_of_peak_sustained_elapsed','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__cycles_elapsed.max','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum','lts__t_sectors_srcunit_tex_op_read.sum','sm__inst_executed_pipe_lsu.sum']
for f in sys.argv[1:]:
    out=subprocess.run(['ncu','-i',f,'--page','raw','--csv'],stdout=subprocess.PIPE,text=True).stdout
    rows=list(csv.reader(out.splitlines()))
    hdr,units=rows[0],rows[1]
    for vals in rows[2:]:
        print('==',f, vals[hdr.index('Kernel Name')][:60] if 'Kernel Name' in hdr else '')
        for i,h in enumerate(hdr):
            if h in want: print('  %-82s %-10s %s'%(h,units[i],vals[i]))
