#!/usr/bin/env python
"""Key metrics of the first kernel of an `ncu --set full` capture, one per line (what profiles/*_full_metrics.txt hold).
usage: ncu -i rep.ncu-rep --page raw --csv > raw.csv
       python tools/ncu_metrics.py raw.csv [extra_metric ...]"""
import csv
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_l1tex2xbar_write_bytes.sum', 'sm__cycles_elapsed.max']


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, vals = rows[0], rows[1], rows[2]
    for w in WANT + sys.argv[2:]:
        for i, h in enumerate(hdr):
            if h == w:
                print('%-70s %-16s %s' % (w, units[i], vals[i]))


if __name__ == '__main__':
    main()
