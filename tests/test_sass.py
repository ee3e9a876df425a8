"""
What the built kernels are made of (no GPU needed: ``cuobjdump -sass`` over the objects of the in-tree build):
sm_100a cubins, TMA bulk copies and mbarriers in the row and chain kernels, no floating-point atomics or reductions
anywhere (every sum has a fixed order), MUFU approximations in the fp32 kernels only.  Skipped where the objects or
cuobjdump are missing.
"""
import glob
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, 'sakkara_b200', 'csrc', 'build')
CUOBJDUMP = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'


def _objects():
    return sorted(glob.glob(os.path.join(BUILD, '*.o')))


pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP) or not _objects(),
                                reason='needs cuobjdump and the objects of the in-tree build')


def _sass(path):
    return subprocess.run([CUOBJDUMP, '-sass', path], capture_output=True, text=True, check=True).stdout


@pytest.fixture(scope='module')
def listings():
    # one row-kernel object per floating type is enough for the mnemonic checks (each is ~200k lines)
    wanted = ['skk_api.o', 'skk_chain_f32.o', 'skk_chain_f64.o', 'skk_row_f32_poisson_bt1.o', 'skk_row_f64_normal_bt4.o']
    return {name: _sass(os.path.join(BUILD, name)) for name in wanted if os.path.exists(os.path.join(BUILD, name))}


def test_every_cubin_is_sm_100a(listings):
    for name, text in listings.items():
        archs = set(re.findall(r'arch = (sm_\w+)', text))
        assert archs == {'sm_100a'}, (name, archs)


def test_no_floating_point_atomics(listings):
    for name, text in listings.items():
        bad = re.findall(r'\b(?:ATOM\w*|RED\w*)\.[\w.]*F(?:32|64)[\w.]*', text)
        assert not bad, (name, sorted(set(bad))[:5])


def test_row_and_chain_kernels_use_tma_and_mbarriers(listings):
    for name, text in listings.items():
        if name == 'skk_api.o':
            continue
        assert 'UBLKCP.S.G' in text, name                       # cp.async.bulk (1-D TMA) into shared memory
        assert 'SYNCS.ARRIVE.TRANS64' in text and 'SYNCS.PHASECHK.TRANS64.TRYWAIT' in text, name
    assert 'ELECT' in listings['skk_row_f32_poisson_bt1.o']     # one elected lane issues the copies


def test_fp32_kernels_use_hardware_approximations_fp64_do_not(listings):
    assert 'MUFU.EX2' in listings['skk_row_f32_poisson_bt1.o']
    assert 'MUFU.EX2' not in listings['skk_row_f64_normal_bt4.o']
