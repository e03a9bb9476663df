"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`, the oracle's
literal numpy/FFT path on the host cores) prints exactly ONE JSON line on stdout with the keys the
driver reads, and the N>1 ranks other than 0 exit quietly."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                           '--steps', '1', '--warmup', '0', '--cpu-sweeps', '2'],
                          capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    p = _run({})
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'logL evals/s' and d['unit'] == 'evals/s'
    assert d['higher_is_better'] is True and d['value'] > 0 and d['vs_baseline'] is None
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['cpu_baseline']['value'] == d['value'] == d['e2e']['value']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config']


def test_reference_arm_other_ranks_exit_quietly():
    p = _run({'RANK': '1', 'LOCAL_RANK': '1', 'WORLD_SIZE': '2'})
    assert p.returncode == 0, p.stderr[-2000:]
    assert p.stdout.strip() == ''
