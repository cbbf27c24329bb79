"""Summarise an .ncu-rep (raw page) into a small text file for profiles/.  Usage: ncu_summary.py rep out title"""
import csv
import subprocess
import sys

rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__ops_path_tensor_src_fp64.sum',
        'sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__cycles_active.avg', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_op_shfl.sum' ]
with open(out, 'w') as f:
    f.write(title + '\n')
    for vals in rows[2:]:
        f.write('\n')
        for k in keep:
            if k in hdr:
                i = hdr.index(k)
                f.write(f'{k} = {vals[i]} {units[i]}\n')
        f.write('warp stall reasons per issue (smsp__average_warps_issue_stalled_*_per_issue_active.ratio):\n')
        st = [(float(vals[i].replace(',', '')), h) for i, h in enumerate(hdr)
              if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and vals[i]]
        for v, h in sorted(st, reverse=True)[:8]:
            f.write(f'  {h.split("issue_stalled_")[1].split("_per_issue")[0]} = {v}\n')
print(open(out).read())
