"""
The C-ABI library loads on a machine without a GPU and exports every symbol that
include/skk_b200.h declares; descriptors are validated before any CUDA call; without a device the
library fails loudly (there is no CPU fallback).
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

from sakkara_b200 import build_ext, runtime

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include', 'skk_b200.h')


@pytest.fixture(scope='module')
def lib():
    if not os.path.exists(build_ext.LIB_PATH):
        build_ext.build()
    return runtime.load_library()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(skk_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_exported(lib):
    names = declared_symbols()
    assert set(names) == set(runtime.EXPORTS), (names, runtime.EXPORTS)
    for name in names:
        assert hasattr(lib, name), f'{name} is declared in skk_b200.h but not exported'


def test_abi_version_matches_header(lib):
    version = int(re.search(r'#define SKK_ABI_VERSION (\d+)', open(HEADER).read()).group(1))
    assert lib.skk_abi_version() == version == runtime.ABI_VERSION


def test_struct_sizes_match_the_header_layout():
    # natural alignment of the C structs in include/skk_b200.h
    assert C.sizeof(runtime.TermDesc) == 24
    assert C.sizeof(runtime.BlockDesc) == 72
    assert C.sizeof(runtime.PlanDesc) == 168
    assert C.sizeof(runtime.Stats) == 80


def test_invalid_descriptors_are_rejected_with_a_message(lib):
    handle = C.c_void_p()
    desc = runtime.PlanDesc(abi_version=99)
    assert lib.skk_plan_create(C.byref(desc), C.byref(handle)) == -1
    assert b'ABI' in lib.skk_last_error()
    desc = runtime.PlanDesc(abi_version=runtime.ABI_VERSION, dtype=7)
    assert lib.skk_plan_create(C.byref(desc), C.byref(handle)) == -1
    assert b'dtype' in lib.skk_last_error()
    assert lib.skk_plan_create(None, C.byref(handle)) == -1
    assert lib.skk_eval(None, None, 1, None, None, 0) == -1
    assert lib.skk_plan_destroy(None) == 0


def test_no_cpu_fallback_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    from sakkara_b200 import synth
    from sakkara_b200.valuegrad import ValueGradFunction
    f = ValueGradFunction(synth.linreg_plan(synth.linreg_columns(100, 3)[0]))
    with pytest.raises(runtime.SkkError, match='no CUDA device|CUDA'):
        f(np.zeros(f.size))


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(RuntimeError, match='not built'):
        runtime.load_library(str(tmp_path / 'libskk_b200.so'))
