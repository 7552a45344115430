"""CPU: the C-ABI shared library loads and exports every symbol include/pyprob_b200.h declares."""
import ctypes
import os
import re

from pyprob_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'pyprob_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ppb_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_something():
    syms = _declared_symbols()
    assert len(syms) >= 30
    assert 'ppb_ic_loss_forward' in syms and 'ppb_weights_finalize' in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), 'build first: python -c "import __graft_entry__ as g; g.build()"'
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, 'declared in include/pyprob_b200.h but not exported: {}'.format(missing)


def test_binding_table_matches_header():
    declared = set(_declared_symbols())
    bound = set(_lib.EXPORTED_SYMBOLS)
    assert declared == bound, (sorted(declared - bound), sorted(bound - declared))


def test_version_without_gpu():
    assert _lib.call('ppb_version') >= 100
