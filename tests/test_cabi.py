"""CPU: the C-ABI library cross-compiles for sm_100a, loads, and exports every symbol include/eyoc_b200.h declares.
No compute call is made (there is no GPU here)."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, 'include', 'eyoc_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(eyoc_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported(built_lib):
    lib = ctypes.CDLL(built_lib)
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/eyoc_b200.h but not exported'
    lib.eyoc_version.restype = ctypes.c_int
    assert lib.eyoc_version() >= 100


def test_sm100a_sass_only(built_lib):
    out = subprocess.run(['cuobjdump', '-lelf', built_lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_\d+a?', out))
    assert archs == {'sm_100a'}, archs


def test_argument_errors_without_gpu(built_lib):
    lib = ctypes.CDLL(built_lib)
    lib.eyoc_last_error.restype = ctypes.c_char_p
    rc = lib.eyoc_knn1(None, None, 1, ctypes.c_int64(4), ctypes.c_int64(4), 32, 0, None, ctypes.c_size_t(0), None, None, None)
    assert rc == -1 and b'null' in lib.eyoc_last_error()
    rc = lib.eyoc_sparse_conv(None, 1, None, 0, None, 1, ctypes.c_int64(1), None, None, None, None, None, 0, 0, None, 1, None)
    assert rc == -1


def test_product_path_fails_loudly_on_cpu_tensors():
    import pytest
    import torch
    from eyoc_b200.lib.eval import find_nn_gpu
    from eyoc_b200.scripts.SC2_PCR.SC2_PCR import Matcher
    with pytest.raises(RuntimeError):
        find_nn_gpu(torch.zeros(4, 32), torch.zeros(4, 32))
    with pytest.raises(RuntimeError):
        Matcher().SC2_PCR(torch.zeros(1, 50, 3), torch.zeros(1, 50, 3))


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'eyoc_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f'{f} imports the oracle'
