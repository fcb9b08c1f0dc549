"""the C-ABI library loads and exports every symbol include/fdts.h declares (no compute calls:
this runs without a GPU), and the product path fails loudly without a CUDA device."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "fdts.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fdts_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from firedrake_ts_b200.engine import LIB_PATH, SYMBOLS

    assert os.path.exists(LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fdts.h but not exported"
    for n in SYMBOLS:
        assert n in names, f"{n} bound by engine.py but not declared in include/fdts.h"


def test_version_and_error_strings():
    from firedrake_ts_b200.engine import load_library

    lib = load_library()
    assert b"sm_100a" in lib.fdts_version()
    assert isinstance(lib.fdts_last_error(), bytes)


def test_null_arguments_are_rejected_without_gpu():
    from firedrake_ts_b200.engine import load_library

    lib = load_library()
    assert lib.fdts_create(None, None) != 0
    assert b"null" in lib.fdts_last_error()
    assert lib.fdts_destroy(None) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_engine_fails_loudly_without_gpu():
    from firedrake_ts_b200.engine import Engine, EngineError
    from firedrake_ts_b200.workloads import make_problem

    form, bcs, u, udot = make_problem("heat", 2, 1, 4)
    with pytest.raises(EngineError, match="no CPU fallback"):
        Engine(form, bcs)


def test_product_package_does_not_import_oracle():
    """the oracle is test infrastructure: nothing under firedrake_ts_b200/ may import it."""
    pkg = os.path.join(ROOT, "firedrake_ts_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
