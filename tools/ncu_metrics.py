"""Print selected metrics of the first kernel in an ncu report: python tools/ncu_metrics.py report.ncu-rep [substr ...]"""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__registers_per_thread',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu.sum', 'lts__t_bytes.sum']
rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''
    print('==', name[:100])
    for i, h in enumerate(hdr):
        if h in KEYS or any(e in h for e in extra):
            print('  %-90s %s %s' % (h, r[i], units[i]))
