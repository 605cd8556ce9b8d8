"""Summarise an `ncu --page raw --csv` dump: key throughput counters and top stall reasons."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'launch__waves_per_multiprocessor', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum',
        'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'sm__inst_executed_pipe_lsu.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    print('=== %s  grid=%s block=%s' % (name[:70], r[hdr.index('Grid Size')], r[hdr.index('Block Size')]))
    for w in KEYS:
        if w in hdr:
            print('  %-66s %16s %s' % (w, r[hdr.index(w)], units[hdr.index(w)]))
    stalls = []
    for i, h in enumerate(hdr):
        if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
            try: stalls.append((float(r[i].replace(',', '')), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
            except ValueError: pass
    for v, h in sorted(stalls, reverse=True)[:8]:
        print('  stall %-40s %.2f warps/issue' % (h, v))
