#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` dump: the metrics DESIGN.md / profiles/ quote, per launch."""
import csv
import sys

KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_blocks', 'launch__occupancy_limit_warps',
        'launch__waves_per_multiprocessor', 'launch__grid_size', 'launch__block_size', 'sm__cycles_elapsed.avg', 'sm__inst_executed.sum', 'smsp__inst_executed.avg.per_cycle_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_xu.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_fp64.sum', 'smsp__thread_inst_executed.sum', 'sm__sass_thread_inst_executed_op_ffma_pred_on.sum',
        'sm__sass_thread_inst_executed_op_fadd_pred_on.sum', 'sm__sass_thread_inst_executed_op_fmul_pred_on.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'inst_executed', 'thread_inst_executed_true', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed', 'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed', 'sm__sass_thread_inst_executed_op_ffma_pred_on.sum.peak_sustained']


def main(path):
    rows = list(csv.reader(open(path)))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        print('-' * 100)
        for k in KEYS:
            if k in h:
                i = h.index(k)
                print(f'{k} = {r[i][:110]} {units[i]}')
        st = [(h[i], r[i]) for i in range(len(h)) if 'smsp__average_warp' in h[i] and 'issue_stalled' in h[i] and h[i].endswith('.ratio') and 'not_issued' not in h[i]]
        st = sorted(st, key=lambda t: -float(t[1].replace(',', '') or 0))
        print('top stall reasons (warp latency cycles per issued instruction):')
        for k, v in st[:8]:
            print('   ', k.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', '').replace('.ratio', ''), v)


if __name__ == '__main__':
    main(sys.argv[1])
