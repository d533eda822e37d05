"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports
every symbol include/evoke_b200.h declares, with the arity the ctypes table assumes."""
import ctypes
import os
import re

import pytest

from evoke_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "evoke_b200.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"typedef struct.*?\}\s*\w+;", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"EVK_API\s+[\w\s\*]+?\b(evk_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[m.group(1)] = n
    return out


def test_library_is_built_and_loads_without_gpu():
    assert os.path.isfile(_lib.LIB_PATH), "run `python -m evoke_b200.build` (or __graft_entry__.build())"
    lib = _lib.load()
    assert lib.evk_version() == _lib.ABI_VERSION
    assert _lib.last_error() == ""


def test_every_declared_symbol_is_exported_with_matching_arity():
    decl = _declared()
    assert len(decl) >= 14
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name, nargs in decl.items():
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} missing from the ctypes table"
        assert len(_lib.SIGNATURES[name]) == nargs, f"{name}: header has {nargs} args, ctypes table {len(_lib.SIGNATURES[name])}"
    assert set(_lib.SIGNATURES) == set(decl)


def test_argument_validation_needs_no_gpu():
    # null pointers / bad shapes are rejected before any CUDA call
    lib = _lib.load()
    rc = lib.evk_l2norm_fwd(None, 0, 4, 8, 8, 1, None, None, 8, None, None, 8, None, None)
    assert rc == _lib.EVK_ERR_INVALID and "non-null" in _lib.last_error()
    with pytest.raises(ValueError):
        _lib.call("evk_posmask_build", None, None, 4, None, None, 4, 0, 0, None, 1, None, None, 0, 0, None, None)
