"""Print the handful of ncu metrics the roofline discussion needs from a .ncu-rep (first kernel in the report, or all).
usage: python tools/ncu_brief.py file.ncu-rep"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_warps', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.max',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    print('==', vals[hdr.index('Kernel Name')][:90])
    for i, h in enumerate(hdr):
        if h in WANT:
            print(f'  {h:75s} {vals[i]} {units[i]}')
        elif 'issue_stalled' in h and 'per_issue_active' in h:
            try:
                if float(vals[i]) >= 0.4:
                    print(f'  stall {h.split("issue_stalled_")[1].split("_per_issue")[0]:30s} {float(vals[i]):.2f}')
            except ValueError:
                pass
