#!/usr/bin/env python3
"""Condense an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of numbers the design
discussion uses: duration, FP64-pipe / issue utilisation, occupancy, DRAM traffic, stall reasons."""
import csv
import io
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__registers_per_thread', 'regs/thread'),
    ('launch__block_size', 'block'), ('launch__grid_size', 'grid'),
    ('smsp__warps_active.avg.per_cycle_active', 'warps/SMSP'),
    ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'fp64 pipe active %'),
    ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'fp64 inst % of peak'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
    ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'xu (MUFU) %'),
    ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lsu %'),
    ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'alu %'),
    ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'fma %'),
    ('sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active', 'uniform %'),
    ('smsp__inst_executed.sum', 'warp instr executed'),
    ('smsp__inst_executed_pipe_fp64.sum', 'fp64 warp instr'),
    ('dram__bytes_read.sum', 'dram read'), ('dram__bytes_write.sum', 'dram write'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'),
    ('lts__t_bytes.sum', 'L2 bytes'),
    ('l1tex__t_sector_hit_rate.pct', 'L1 hit %'),
    ('sm__icc_request_hit_rate.pct', 'icache hit %'),
    ('smsp__inst_executed_op_local_ld.sum', 'local loads'), ('smsp__inst_executed_op_local_st.sum', 'local stores'),
    ('smsp__inst_executed_op_shared_ld.sum', 'shared loads'), ('smsp__inst_executed_op_shared_st.sum', 'shared stores'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem bank conflicts'),
]


def main(path, pattern=None):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if pattern and pattern not in d['Kernel Name']:
            continue
        print(f"== {d['Kernel Name']}  (id {d.get('ID')})")
        for k, label in KEYS:
            if k in d and d[k] != '':
                print(f'  {label:24s} {d[k]} {units[hdr.index(k)]}')
        stalls = []
        for k in hdr:
            if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio'):
                try:
                    stalls.append((float(d[k]), k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print('  stalls per issue:', ', '.join(f'{n}={v:.2f}' for v, n in stalls[:8]))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
