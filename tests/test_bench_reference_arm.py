"""bench.py --impl reference (the driver's CPU arm) runs without a GPU: one JSON line with the contract's keys, and under
torch.distributed.run only rank 0 prints while every rank exits 0."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
        'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e')


def check_line(out: str, n_gpus: int):
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in KEYS:
        assert k in d, k
    assert d['impl'] == 'reference' and d['n_gpus'] == n_gpus and d['value'] > 0
    assert d['unit'] == 'sentences/s' and d['higher_is_better'] is True and d['vs_baseline'] is None
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['cpu_baseline']['value'] == d['value'] == d['e2e']['value']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_reference_arm_single_process():
    p = subprocess.run([sys.executable, 'bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1',
                        '--cpu-sample', '2'], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    check_line(p.stdout, 1)


def test_reference_arm_under_torchrun_world_2():
    p = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', '29537', 'bench.py', '--impl', 'reference',
                        '--gpus', '2', '--steps', '1', '--warmup', '1', '--cpu-sample', '2'],
                       cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    check_line(p.stdout, 2)
