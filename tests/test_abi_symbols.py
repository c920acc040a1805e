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
    assert _lib.lib().ptb_abi_version() == 5


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
    try:
        engine.MultiContext([0, 1])
    except PtbError as e:
        assert e.code == -2 and "no CPU fallback" in e.message
    else:
        raise AssertionError("ptb_multi_create succeeded without a GPU")


def test_product_does_not_import_oracle():
    """The product package must never reach into oracle/."""
    pkg = ROOT / "path_trace_golang_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.h")):
        text = p.read_text().lower()
        for needle in ("import oracle", "from oracle", "liboracle", "oracle/", "pyoracle"):
            assert needle not in text, (p, needle)


def test_headers_are_plain_c():
    """include/*.h must be consumable by cgo: they compile as C11 (no C++ in the signatures)."""
    import subprocess
    for h in sorted((ROOT / "include").glob("*.h")):
        r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", str(h)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_bench_reference_arm_schema():
    """bench.py --impl reference (the CPU arm the driver runs first): one JSON line with the contract's keys; only rank 0 prints."""
    import json
    import os
    import subprocess
    import sys
    cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "C1", "--steps", "1", "--warmup", "0", "--cpu-spp", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ["impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"]:
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "Msamples/s" and line["value"] > 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and "workload" in line["config"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=120, env={**os.environ, "RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cgo_shim_only_uses_declared_entry_points():
    """go/internal/engine/cuda/cuda.go cannot be compiled here (no Go toolchain): at least every C.ptb_* it calls must be
    declared in include/ptb200.h and exported by the library."""
    import re
    from path_trace_golang_b200 import _lib
    src = (ROOT / "go" / "internal" / "engine" / "cuda" / "cuda.go").read_text()
    header = (ROOT / "include" / "ptb200.h").read_text()
    used = set(re.findall(r"C\.(ptb_[a-z0-9_]+)\(", src))
    assert {"ptb_create", "ptb_scene_upload", "ptb_render", "ptb_multi_create", "ptb_multi_render"} <= used
    lib = _lib.lib()
    for name in used - {"ptb_go_progress_ptr"}:            # (the trampoline getter is defined in the cgo preamble)
        assert re.search(r"\b%s\s*\(" % name, header), f"{name} is not declared in ptb200.h"
        assert hasattr(lib, name), f"{name} is not exported"
    # struct fields the shim fills must exist in the header's ptb_scene / ptb_cfg
    for field in re.findall(r"\bf\.s\.([a-z_0-9]+)\s*=", src):
        assert re.search(r"\b%s\b" % field, header), f"ptb_scene.{field} missing from the header"
