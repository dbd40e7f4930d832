"""Turn the scratch outputs of tools/profile_round.sh (gpurun_out/) into the tracked summaries under profiles/.

    python tools/summarize_profiles.py r01b

Writes profiles/launches_<R>_summary.csv (per-kernel share of device time from the ncu launch list),
profiles/ncu_gemm_<R>.csv (one row per captured tcgen05 GEMM launch: duration, DRAM bytes, tensor-pipe and
memory-throughput counters from `ncu --set full`), profiles/traffic_<R>.json (what bench.py reports as
roofline.traffic) and copies the bench JSON lines.
"""
import csv
import json
import os
import re
import shutil
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, 'gpurun_out')
P = os.path.join(ROOT, 'profiles')


def short(name):
    m = re.search(r'(\w+)(<[^>]*>)?\(', name)
    n = (m.group(1) + (m.group(2) or '')) if m else name
    return n.replace('(int)', '').replace('(bool)', '')


def launches(R):
    src = os.path.join(G, f'launches_{R}.csv')
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        a = agg[short(r[ki])]
        a[0] += 1
        a[1] += float(r[vi].replace(',', '')) / 1e3
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, f'launches_{R}_summary.csv'), 'w') as f:
        f.write(f'# ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e` (default batch, L=1)\n')
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES\n')
        f.write(f'# {sum(v[0] for v in agg.values())} launches, total {tot / 1e3:.3f} ms\n')
        f.write('kernel,launches,total_us,share\n')
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'"{k}",{v[0]},{v[1]:.1f},{v[1] / tot:.4f}\n')
    shutil.copy(src, os.path.join(P, f'launches_{R}.csv'))


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'launch__registers_per_thread', 'smsp__inst_executed.sum']


LN_FUSED = (8, 10, 12, 14)


def read_raw(name):
    """`ncu --page raw --csv` of gpurun_out/<name>: exported on the GPU box (<name>.raw.csv; reports with sources exceed
    what gpurun brings back) or from the report itself when it is here."""
    pre = os.path.join(G, name + '.raw.csv')
    if os.path.isfile(pre):
        return open(pre).read()
    rep = os.path.join(G, name + '.ncu-rep')
    return subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout


def gemm(R):
    raw = read_raw(f'gemm_{R}')
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    tensor_cols = [h for h in hdr if 'pipe_tensor' in h and 'pct' in h]
    keep = [w for w in WANT if w in col] + [h for h in tensor_cols if h not in WANT]
    out_rows, traffic = [], []
    for r in rows[2:]:
        name = short(r[col['Kernel Name']])
        rec = {'kernel': name, 'grid': r[col['launch__grid_size']]}
        for k in keep:
            rec[k + ' [' + units[col[k]] + ']'] = r[col[k]]

        def val(k, scale):
            u = units[col[k]].lower()
            v = float(r[col[k]].replace(',', ''))
            mult = {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(u, 1)
            return v * mult / scale
        rec['dram_bytes'] = val('dram__bytes_read.sum', 1) + val('dram__bytes_write.sum', 1)
        rec['duration_us'] = val('gpu__time_duration.sum', 1)
        out_rows.append(rec)
        traffic.append(rec['dram_bytes'])
    # bench.py's `linear_bf16_tcgen05` row (roofline.achieved) holds the plain GEMM ops only; the four GEMMs of a step whose
    # split-K partials go through the fused reduce + LayerNorm kernel (ops.linear_ln: launches 8, 10, 12, 14 of the 15, the
    # single-query encoders' output projections) are reported in their own row, so they are left out of the traffic mean too
    for i, rec in enumerate(out_rows):
        rec['in_gemm_roofline_row'] = int(not (len(out_rows) == 15 and i in LN_FUSED))
    traffic = [t for t, rec in zip(traffic, out_rows) if rec['in_gemm_roofline_row']]
    keys = list(out_rows[0].keys())
    try:          # the batch of the bench line captured in the same round
        batch = json.load(open(os.path.join(G, f'bench_{R}.json')))['config']['batch_per_gpu']
    except Exception:
        batch = 1024
    with open(os.path.join(P, f'ncu_gemm_{R}.csv'), 'w') as f:
        f.write('# ncu --set full --clock-control none -k regex:gemm_bf16_tcgen05 --launch-skip 45 -c 15 '
                'python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-widened --no-configs\n')
        f.write(f'# = the 15 tcgen05 GEMM launches of one bench step (B={batch}, L=1), in launch order\n')
        w = csv.DictWriter(f, fieldnames=keys)
        w.writeheader()
        for rec in out_rows:
            w.writerow(rec)
    json.dump({'source': f'profiles/ncu_gemm_{R}.csv', 'workload': f'B{batch}_L1', 'launches': len(traffic),
               'dram_bytes_per_launch_mean': sum(traffic) / len(traffic), 'dram_bytes_per_step': sum(traffic)},
              open(os.path.join(P, f'traffic_{R}.json'), 'w'), indent=1)


def main():
    R = sys.argv[1] if len(sys.argv) > 1 else 'r01'
    os.makedirs(P, exist_ok=True)
    launches(R)
    gemm(R)
    for name in [f'bench_{R}.json', f'bench_ref_{R}.json'] + [f'bench_{R}_{x}.json' for x in (
            'L5', 'hires', 'B256', 'train128', 'train32', 'train32_bilstm')]:
        src = os.path.join(G, name)
        if os.path.isfile(src):
            lines = [l for l in open(src) if l.startswith('{')]
            open(os.path.join(P, name), 'w').write(lines[-1] if lines else '')
    print(open(os.path.join(P, f'launches_{R}_summary.csv')).read())
    print(open(os.path.join(P, f'traffic_{R}.json')).read())


if __name__ == '__main__':
    main()
