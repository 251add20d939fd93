"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`, the CPU port of the reference's
`_ref` generator forward) prints ONE JSON line with the keys the driver reads, on this container's host cores, and the
ranks other than 0 of a multi-process launch stay silent."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    env['CUDA_VISIBLE_DEVICES'] = ''
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_the_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'generator_slices_per_sec_256x256' and d['unit'] == 'slices/s'
    assert d['higher_is_better'] is True and d['scaling'] == 'weak' and d['vs_baseline'] is None and d['data'] == 'synthetic'
    assert d['value'] > 0 and abs(d['value'] - 1e3 / d['ms_per_step']) < 1e-6 * d['value']
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] == os.cpu_count() and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == dict(value=d['value'], unit='slices/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert d['gpu_launches'] == 0 and 'workload' in d['config'] and 'model' not in d['config']


def test_reference_arm_other_ranks_stay_silent():
    assert _run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'}) == []
