"""Print the handful of metrics we care about from `ncu -i X.ncu-rep --page raw --csv` (stdin or file args)."""
import csv, sys, subprocess
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_red.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum',
        'smsp__average_warps_issue_stalled', 'smsp__pcsamp_warps_issue_stalled', 'sm__inst_executed_pipe_lsu', 'sm__pipe_alu_cycles_active',
        'sm__pipe_fma_cycles_active', 'lts__throughput.avg.pct', 'l1tex__throughput.avg.pct', 'sm__cycles_elapsed.avg ',
        'smsp__cycles_active.avg ', 'sm__inst_executed_pipe_', 'lts__t_sector_hit_rate.pct']
for f in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', f, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for data in rows[2:]:
        print('==', f, data[hdr.index('Kernel Name')], 'grid', data[hdr.index('Grid Size')], 'block', data[hdr.index('Block Size')])
        for h, u, v in zip(hdr, units, data):
            hh = h.split('.', 2)[-1] if h.count('.') >= 2 and h.split('.')[1].startswith('Triage') else h
            if any(k in h for k in KEYS):
                print('   %-95s %-12s %s' % (h[-95:], u, v))
