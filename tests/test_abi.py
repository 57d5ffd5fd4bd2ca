"""The C-ABI library loads on a CPU-only box and exports every symbol include/lsnet_b200.h declares; without a
device it fails loudly rather than falling back."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'lsnet_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(lsnet_[a-z0-9_]+)\s*\(', src)))


def test_all_declared_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/lsnet_b200.h but not exported'


def test_fails_loudly_without_device(lib):
    import torch
    if torch.cuda.is_available():
        return
    lib.lsnet_last_error.restype = ctypes.c_char_p
    assert lib.lsnet_require_sm100() != 0
    assert b'CUDA' in lib.lsnet_last_error() or b'device' in lib.lsnet_last_error()
