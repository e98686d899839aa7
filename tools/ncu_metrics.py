"""Print the headline counters of an `ncu --page raw --csv` export (development aid)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1]))); hdr = rows[0]; units = rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warp_latency_per_inst_issued.ratio'] + \
       ['smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % k for k in
        ('long_scoreboard', 'wait', 'short_scoreboard', 'mio_throttle', 'no_instruction', 'barrier', 'branch_resolving', 'lg_throttle', 'math_pipe_throttle', 'not_selected', 'dispatch_stall', 'membar', 'sleeping', 'drain')] + \
       ['smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum', 'lts__t_bytes_srcunit_tex_aperture_device_op_write.sum']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print('| %s | %s | %s |' % (w, units[i], ' | '.join(r[i] for r in rows[2:])))
