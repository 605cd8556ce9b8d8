"""Summarise an `ncu --page raw --csv` dump: the metrics the roofline notes use + top stall reasons."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'lts__t_sectors_srcunit_tex_op_write.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'smsp__inst_executed_op_shared_atom.sum', 'sm__cycles_elapsed.max', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('-----')
    for w in want:
        if w in d:
            print(' ', w, d[w], units[hdr.index(w)])
    st = [(h, d[h]) for h in hdr if 'issue_stalled' in h and h.endswith('.ratio') and 'not_issued' not in h]
    st = [(h, float(v.replace(',', ''))) for h, v in st if v]
    st.sort(key=lambda x: -x[1])
    for h, v in st[:8]:
        print('     stall', h.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', ''), round(v, 2))
