"""CPU tests of bench.py's contract for the reference arm (`--impl reference`): it runs the reference's own generated
SERIAL code (oracle/_ref) on the host cores, so it needs no GPU -- one JSON line with the agreed keys; under a
multi-rank launch only rank 0 works and prints."""
import json
import os
import subprocess
import sys

import pytest

from tests.common import ROOT


def _run(extra_env=None, *args):
    env = dict(os.environ)
    for k in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE'):
        env.pop(k, None)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                           '--warmup', '1', *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                          env=env, timeout=600)


def test_reference_arm_prints_one_json_line():
    if not os.path.exists(os.path.join(ROOT, 'oracle', '_ref')):
        pytest.skip('oracle/_ref not built')
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, 'only the JSON line may reach stdout'
    rec = json.loads(lines[0])
    assert rec['impl'] == 'reference' and rec['unit'] == 'states/s' and rec['higher_is_better'] is True
    assert rec['metric'].startswith('BK1+BK2 states/sec') and rec['dtype'] == 'f64' and rec['n_gpus'] == 1
    assert rec['steps'] == 1 and rec['warmup'] == 1 and rec['value'] > 0 and rec['ms_per_step'] > 0
    assert 'workload' in rec['config'] and 'GRI-Mech 3.0' in rec['config']['workload']
    cb = rec['cpu_baseline']
    assert cb['kind'] in ('reference', 'port') and cb['cores'] >= 1 and cb['value'] == rec['value'] and cb['sample']
    e2e = rec['e2e']
    assert e2e['value'] == rec['value'] and e2e['h2d_bytes_per_step'] == 0 and e2e['d2h_bytes_per_step'] == 0
    assert rec['gpu_launches'] == 0 and rec['vs_baseline'] is None


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({'RANK': '1', 'LOCAL_RANK': '1', 'WORLD_SIZE': '2'}, '--gpus', '2')
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ''
