"""
bench.py on a box without a GPU: the CPU arm (`--impl reference`) prints one JSON line with the keys of
the contract, and the product arm refuses to run (there is no CPU fallback on the product path).
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), *args], capture_output=True, text=True,
                          timeout=600, env={**os.environ, **(env or {})})


@pytest.mark.parametrize('config', ['c2', 'c3', 'c4'])
def test_reference_arm_line(config):
    proc = run('--impl', 'reference', '--config', config, '--steps', '2', '--warmup', '1', '--ref-rows', '30000')
    assert proc.returncode == 0, proc.stderr
    line = json.loads(proc.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['metric'] == 'logp_dlogp_row_evals_per_s'
    assert line['unit'] == 'row-evals/s' and line['higher_is_better'] is True and line['value'] > 0
    assert line['steps'] == 2 and line['warmup'] == 1 and line['n_gpus'] == 1
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['cpu_baseline']['value'] == line['value']
    assert line['e2e'] == {'value': line['value'], 'unit': 'row-evals/s', 'h2d_bytes_per_step': 0,
                           'd2h_bytes_per_step': 0}
    assert line['dtype'] in ('f32', 'f64') and line['data'] == 'synthetic' and 'workload' in line['config']


def test_reference_arm_other_ranks_are_silent():
    proc = run('--impl', 'reference', '--config', 'c2', '--steps', '1', '--warmup', '1', '--ref-rows', '1000',
               env={'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert proc.returncode == 0 and proc.stdout.strip() == ''


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    proc = run('--config', 'c2', '--steps', '1', '--warmup', '1', '--rows', '1000')
    assert proc.returncode != 0
    assert 'no CPU fallback' in (proc.stderr + proc.stdout)
