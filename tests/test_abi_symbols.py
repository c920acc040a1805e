"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares."""
import ctypes
import re

from conftest import ROOT


def declared_symbols():
    names = set()
    for h in (ROOT / "include").glob("*.h"):
        text = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        for m in re.finditer(r"\b(ptb_[a-z0-9_]+)\s*\(", text):
            names.add(m.group(1))
    names -= {"ptb_progress_fn"}
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from path_trace_golang_b200 import _lib
    L = ctypes.CDLL(str(_lib.SO_PATH))
    decl = declared_symbols()
    assert len(decl) >= 25
    missing = [n for n in decl if not hasattr(L, n)]
    assert not missing, missing
    # and the ctypes table binds exactly the declared set
    assert sorted(_lib.SYMBOLS) == decl


def test_abi_version():
    from path_trace_golang_b200 import _lib
    assert _lib.lib().ptb_abi_version() == 2


def test_create_fails_loudly_without_gpu():
    """No CPU fallback: on a box without CUDA the context cannot be created and says why."""
    import torch
    from path_trace_golang_b200 import engine, PtbError
    if torch.cuda.is_available():
        return
    try:
        engine.Context(0)
    except PtbError as e:
        assert e.code == -2 and "no CPU fallback" in e.message
    else:
        raise AssertionError("ptb_create succeeded without a GPU")


def test_product_does_not_import_oracle():
    """The product package must never reach into oracle/."""
    pkg = ROOT / "path_trace_golang_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.h")):
        text = p.read_text().lower()
        for needle in ("import oracle", "from oracle", "liboracle", "oracle/", "pyoracle"):
            assert needle not in text, (p, needle)
