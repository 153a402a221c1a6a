"""Brief per-kernel table from `ncu -i X.ncu-rep --page raw --csv`: time, issue rate, occupancy, DRAM, stall reasons."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
base = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'smsp__inst_executed.avg', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('----', r[idx['Kernel Name']][:60])
    for w in base:
        if w in idx:
            print('  %-70s %s %s' % (w, r[idx[w]], units[idx[w]]))
    st = [(float(r[i] or 0), h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')) for h, i in idx.items()
          if 'issue_stalled_' in h and h.endswith('_per_issue_active.ratio') and 'average_warps' in h]
    print('  stalls per issue:', ', '.join('%s %.2f' % (n, v) for v, n in sorted(st, reverse=True)[:8]))
