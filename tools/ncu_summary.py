#!/usr/bin/env python3
"""Print the headline metrics of every kernel in an `ncu --page raw --csv` dump."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__block_size', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_blocks',
        'sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
        'smsp__inst_executed_op_global_ld.sum', 'smsp__inst_executed_op_global_st.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum']
for r in rows[2:]:
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(w, '[%s]' % units[i], '=', r[i])
    for i, h in enumerate(hdr):
        if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio'):
            try:
                v = float(r[i])
                if v > 0.1:
                    print('  stall', h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), round(v, 3))
            except ValueError:
                pass
    print()
