"""The C-ABI library builds, loads and exports every symbol include/xmeta.h declares (no compute calls:
this runs without a GPU), and the ctypes mirror of the argument blocks has the C layout."""
import ctypes
import os
import re
import subprocess

import pytest

from exploring_meta_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'xmeta.h')


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(xm_[a-z0-9_]+)\s*\(', src)))


@pytest.fixture(scope='module')
def lib():
    build.build()
    return _lib.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = _declared_functions()
    assert len(names) >= 15
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), 'libxmeta.so does not export %s' % n
        assert n in _lib.SYMBOLS, '_lib.SYMBOLS has no binding for %s' % n
    assert sorted(_lib.SYMBOLS) == names


def test_version_and_error_slot(lib):
    assert lib.xm_version() == 100
    assert lib.xm_last_error() is not None
    # invalid geometry is rejected on the host before any CUDA call
    a = _lib.XmConvArgs()
    assert lib.xm_conv(ctypes.byref(a), None) < 0
    assert b'geom' in lib.xm_last_error().lower() or len(lib.xm_last_error()) > 0


def test_struct_layout_matches_c(tmp_path):
    """sizeof/offsetof of every argument block as the C compiler sees them == the ctypes mirror."""
    structs = {'XmBlockGeom': _lib.XmBlockGeom, 'XmConvArgs': _lib.XmConvArgs, 'XmWgradArgs': _lib.XmWgradArgs,
               'XmBnArgs': _lib.XmBnArgs, 'XmImgArgs': _lib.XmImgArgs, 'XmHeadArgs': _lib.XmHeadArgs,
               'XmAnilHeadArgs': _lib.XmAnilHeadArgs, 'XmSampleArgs': _lib.XmSampleArgs,
               'XmAdamArgs': _lib.XmAdamArgs, 'XmRlAdvArgs': _lib.XmRlAdvArgs, 'XmRlSweepArgs': _lib.XmRlSweepArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "xmeta.h"', 'int main(void){']
    for name, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (name, name))
        for f, _t in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, f, name, f))
    lines.append('return 0;}')
    c = tmp_path / 'layout.c'
    c.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(c), '-o', str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for name, cls in structs.items():
        assert int(out[name]) == ctypes.sizeof(cls), name
        for f, _t in cls._fields_:
            assert int(out['%s.%s' % (name, f)]) == getattr(cls, f).offset, '%s.%s' % (name, f)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(_lib.XmetaError):
        _lib.load()
