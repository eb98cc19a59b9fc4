"""CPU: the C-ABI library builds/loads and exports every symbol the public header declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "zeroshape_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zs_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_symbols():
    syms = _declared_symbols()
    assert len(syms) >= 30 and "zs_chamfer_nn_fwd" in syms and "zs_chain_qkvattn_fwd" in syms


def test_library_exports_every_declared_symbol():
    from zeroshape_b200 import build
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_binding_table_matches_header():
    from zeroshape_b200 import _native
    assert sorted(_native.SIGNATURES) == _declared_symbols()
    assert _native.lib.zs_abi_version() == 1


def test_product_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "zeroshape_b200")):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    import importlib
    import subprocess
    import sys
    code = ("import os; os.environ['ZEROSHAPE_B200_LIB']='/nonexistent/lib.so'\n"
            "try:\n    import zeroshape_b200._native\nexcept ImportError as e:\n    print('OK', e)\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert out.stdout.startswith("OK"), out.stdout + out.stderr
