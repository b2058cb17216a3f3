"""CPU-only checks of the C-ABI boundary: the library loads without a GPU, exports every symbol the header
declares, the ctypes table covers exactly those symbols, and error paths report through cfm_last_error().
No compute entry point is exercised here."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conformer_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "conformer_b200.h")


def header_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"CFM_API\s+[\w\s\*]+?\b(cfm_\w+)\s*\(", text)))


def test_library_builds_and_loads():
    from conformer_b200.build import build
    path = build()
    assert os.path.exists(path)
    assert _lib.lib().cfm_abi_version() == 2


def test_every_header_symbol_is_exported_and_bound():
    syms = header_symbols()
    assert len(syms) >= 25
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (cfm_\w+)", out))
    missing = [s for s in syms if s not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    unbound = [s for s in syms if s not in _lib.SIGNATURES]
    assert not unbound, f"declared in the header but missing from the ctypes table: {unbound}"
    stale = [s for s in _lib.SIGNATURES if s not in syms]
    assert not stale, f"bound in ctypes but not declared in the header: {stale}"


def test_no_torch_types_in_the_abi():
    import re as _re
    code = _re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=_re.S)       # declarations only, comments stripped
    assert "at::" not in code and "Tensor" not in code and "torch" not in code.lower()


def test_error_reporting_without_gpu_work():
    lib = _lib.lib()
    h = C.c_void_p()
    cfg = _lib.Config(80, 1, 64, 4, 31)
    assert lib.cfm_module_create(99, C.byref(cfg), C.byref(h)) != 0
    assert b"unknown module kind" in lib.cfm_last_error()
    assert lib.cfm_module_create(_lib.KIND_FFN, C.byref(cfg), C.byref(h)) == 0
    assert lib.cfm_module_set_param(h, b"x", None, 4) != 0          # null data pointer
    assert lib.cfm_workspace_bytes(h, 1, 10) < 0                    # not finalized
    assert b"not finalized" in lib.cfm_last_error()
    assert lib.cfm_module_destroy(h) == 0
    with pytest.raises(_lib.NativeError):
        _lib.check(lib.cfm_ffn_forward(None, None, None, 1, 1, None, 0, None))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.NativeError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_ctypes_argument_counts_match_the_header():
    """Signature drift between include/conformer_b200.h and the ctypes table is an ABI bug no compiler sees: every
    declaration's parameter count (and whether it returns a 64-bit value) must match its binding."""
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    decls = re.findall(r"CFM_API\s+([\w\s\*]+?)\b(cfm_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S)
    assert len(decls) == len(header_symbols())
    for ret, name, params in decls:
        params = params.strip()
        n = 0 if params in ("", "void") else len([p for p in params.split(",") if p.strip()])
        restype, argtypes = _lib.SIGNATURES[name]
        assert n == len(argtypes), f"{name}: header has {n} parameters, ctypes table {len(argtypes)}"
        if "int64_t" in ret:
            assert restype is C.c_int64, f"{name} returns int64_t in the header"
        if "char" in ret:
            assert restype is C.c_char_p, f"{name} returns a string in the header"
