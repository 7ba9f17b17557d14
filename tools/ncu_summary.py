"""Summarise an .ncu-rep (read with `ncu -i`) into a small text file for profiles/."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_bytes.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_warps', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio']
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
    print('kernel:', name)
    for i, h in enumerate(hdr):
        if h in WANT:
            print('  %-85s %-14s %s' % (h, units[i], r[i]))
