#!/usr/bin/env python3
"""Summarise an .ncu-rep (read on the CPU box): key roofline counters + warp stall reasons."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_bytes.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg']


def main(path, which=None):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for k, d in enumerate(data):
        if which is not None and k != which:
            continue
        print('## launch', k, d[hdr.index('Kernel Name')][:90])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f'  {w:70s} {d[i]:>16s} {units[i]}')
        st = []
        for i, h in enumerate(hdr):
            if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
                try:
                    st.append((float(d[i]), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print('  stalls (warps per issue-active):', ', '.join(f'{n}={v:.2f}' for v, n in st[:7]))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
