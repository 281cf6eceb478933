"""CPU: the reference arm of bench.py (`--impl reference`: the unmodified reference classes from
baseline/_ref when the copy is there, else the CPU restatement; one process per chain) prints exactly
ONE JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS='1')
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '1', '--chains-per-gpu', '2', '--cpu-sample-cells', '300'],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout
    out = json.loads(lines[0])
    assert out['impl'] == 'reference' and out['unit'] == 'chain-steps/s' and out['higher_is_better'] is True
    assert out['metric'].startswith('MCMC steps/sec/chain') and out['value'] > 0
    assert out['steps'] == 1 and out['n_gpus'] == 1 and out['data'] == 'synthetic' and out['vs_baseline'] is None
    cpu = out['cpu_baseline']
    from oracle.ref_shim import reference_available
    assert cpu['kind'] == ('reference' if reference_available() else 'port')
    assert cpu['cores'] == 2 and cpu['value'] == out['value'] and '300 of 100000' in cpu['sample']
    assert out['steps_requested'] == 1
    assert out['e2e'] == dict(value=out['value'], unit='chain-steps/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert 'workload' in out['config'] and 'model' not in out['config']
