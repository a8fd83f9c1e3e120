"""Print the handful of raw ncu metrics the notes in profiles/ quote.
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv; python tools/ncu_summary.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('kernel', d.get('Kernel Name', '')[:70])
    for k in ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
              'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
              'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
              'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
              'lts__t_sector_hit_rate.pct', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
              'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
              'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
              'dram__throughput.avg.pct_of_peak_sustained_elapsed']:
        print('  ', k, d.get(k))
    for k in hdr:
        if 'issue_stalled' in k and 'per_issue_active' in k:
            print('   stall', k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), d[k])
