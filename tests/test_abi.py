"""The C-ABI library loads without a GPU and exports every symbol include/f1l.h declares; the
product never touches the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "f1l.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(f1l_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    from f1tenth_planning_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from f1tenth_planning_b200 import build
        build.build()
    names = _declared()
    assert len(names) >= 30
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "missing export " + n
    assert sorted(_lib.SIGNATURES) == names, set(names) ^ set(_lib.SIGNATURES)
    _lib.lib()


def test_struct_layouts_match_header():
    from f1tenth_planning_b200 import _lib
    from oracle import c_oracle as co
    assert ctypes.sizeof(_lib.Config) == 10 * 4 + 5 * 8 + 7 * 8   # 9 int32 + 4 bytes of padding
    assert ctypes.sizeof(_lib.PlanResult) == 16 + 16 + 8 + 9 * 8
    assert ctypes.sizeof(co.Config) == ctypes.sizeof(_lib.Config)
    cfg = _lib.default_config()
    oc = co.default_config()
    for f, _ in _lib.Config._fields_:
        a, b = getattr(cfg, f), getattr(oc, f)
        if f == "weights":
            assert list(a) == list(b)
        else:
            assert a == b, f


def test_no_device_is_reported_not_hidden():
    import torch
    if torch.cuda.is_available():
        pytest.skip("box has a GPU")
    from f1tenth_planning_b200 import _lib
    from f1tenth_planning_b200.engine import Engine
    with pytest.raises(_lib.F1LError, match="no CUDA device"):
        Engine()
    assert b"invalid" in _lib.lib().f1l_strerror(-1)


def test_product_does_not_reference_the_oracle():
    """no import, link or path of oracle/ anywhere in the product package (comments may name it)"""
    pkg = os.path.join(ROOT, "f1tenth_planning_b200")
    bad = re.compile(r"(^\s*(from|import)\s+oracle\b)|(\boracle[./](c_oracle|build|_build|c)\b)|"
                     r"(f1o_)|(libf1o)|(c_oracle)", re.M)
    n = 0
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                n += 1
                text = open(os.path.join(dirpath, f)).read()
                assert not bad.search(text), (f, bad.search(text).group(0))
    assert n >= 10
