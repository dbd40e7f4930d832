"""Summarise any `ncu --set full` report in gpurun_out/ into a tracked CSV under profiles/.

    python tools/summarize_ncu_rep.py attn5_r01 ncu_attn_r01 "how it was captured"

One row per captured launch: duration, DRAM bytes (read + write), achieved DRAM GB/s (bytes / duration -- under ncu's
replay the duration is cold-cache and serialised), DRAM / SM throughput percentages, issue-slot and tensor-pipe activity.
"""
import csv
import os
import subprocess
import sys

from summarize_profiles import G, P, read_raw, short

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_issued.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.avg.per_cycle_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'smsp__inst_executed.sum']
SCALE = {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'usecond': 1, 'ms': 1e3, 'msecond': 1e3,
         'nsecond': 1e-3, 'second': 1e6}


def main(rep_name, out_name, note):
    raw = read_raw(rep_name)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    keep = [w for w in WANT if w in col]

    def val(r, k):
        return float(r[col[k]].replace(',', '')) * SCALE.get(units[col[k]].lower(), 1)

    with open(os.path.join(P, out_name + '.csv'), 'w') as f:
        f.write(f'# {note}\n# source: gpurun_out/{rep_name}.ncu-rep (ncu --set full --clock-control none); one row per launch\n')
        w = csv.writer(f)
        w.writerow(['kernel', 'duration_us', 'dram_bytes', 'dram_GBps'] + [f'{k} [{units[col[k]]}]' for k in keep])
        for r in rows[2:]:
            us = val(r, 'gpu__time_duration.sum')
            b = val(r, 'dram__bytes_read.sum') + val(r, 'dram__bytes_write.sum')
            w.writerow([short(r[col['Kernel Name']]), f'{us:.2f}', f'{b:.0f}', f'{b / us / 1e3:.1f}'] + [r[col[k]] for k in keep])
    print(open(os.path.join(P, out_name + '.csv')).read()[:3000])


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else '')
