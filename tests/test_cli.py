"""tools/render.py — the headless driver (twin of cmd/render/main.go:46-63)."""
import importlib.util
import pathlib
import sys

import numpy as np
import pytest

from conftest import ROOT, scene_path

spec = importlib.util.spec_from_file_location("render_cli", ROOT / "tools" / "render.py")
render_cli = importlib.util.module_from_spec(spec)
spec.loader.exec_module(render_cli)


def test_flag_names_and_settings_resolution():
    from path_trace_golang_b200 import engine, scene
    a = render_cli.parse([])
    assert (a.scene, a.mode, a.out) == ("scenes/example_simple.json", "preview", "output.png")     # main.go:17-21 defaults
    sc = scene.Load(scene_path("example_simple"))
    preview, final = engine.RenderSettingsForMode("preview"), engine.RenderSettingsForMode("final")
    assert render_cli.resolve_settings(render_cli.parse([]), sc.Settings, preview) == (400, 225, 20, 20)      # headless ignores scene.settings (main.go:52)
    assert render_cli.resolve_settings(render_cli.parse(["-mode", "final"]), sc.Settings, final) == (1920, 1080, 1000, 80)
    assert render_cli.resolve_settings(render_cli.parse(["-settings"]), sc.Settings, preview) == (400, 225, 20, 10)   # scene file: depth 10
    c1 = render_cli.parse(["-width", "640", "-height", "360", "-spp", "16", "-depth", "8"])
    assert render_cli.resolve_settings(c1, sc.Settings, preview) == (640, 360, 16, 8)               # BASELINE config C1 is reachable
    zero = scene.Load(scene_path("metal_glass_room")).Settings                                      # all-zero settings block
    assert render_cli.resolve_settings(render_cli.parse(["-settings"]), zero, preview) == (400, 225, 20, 20)


def test_missing_scene_is_an_error(capsys):
    assert render_cli.main(["-scene", "/nonexistent.json"]) == 1
    assert "load scene" in capsys.readouterr().err


@pytest.mark.gpu
def test_headless_render_writes_png(tmp_path):
    from PIL import Image
    out = tmp_path / "o.png"
    rc = render_cli.main(["-scene", str(scene_path("example_simple")), "-width", "160", "-height", "90", "-spp", "8", "-depth", "8",
                          "-out", str(out)])
    assert rc == 0
    img = np.array(Image.open(out))
    assert img.shape == (90, 160, 4) and (img[..., 3] == 255).all() and img[..., :3].std() > 5


def test_cpp_driver_flags_and_errors():
    """ptb_render = cmd/render in C++ over the host mirror: Go-style flags, loud failure (exit 1) without a GPU or scene."""
    import subprocess
    exe = ROOT / "path_trace_golang_b200" / "ptb_render"
    assert exe.exists(), "build with __graft_entry__.build()"
    r = subprocess.run([str(exe), "-bogus"], capture_output=True, text=True)
    assert r.returncode == 2 and "flag provided but not defined: -bogus" in r.stderr
    r = subprocess.run([str(exe), "-scene", "/nonexistent.json", "-headless"], capture_output=True, text=True)
    assert r.returncode == 1 and "open scene" in r.stderr
    # boolean flags follow Go's flag package: -flag=false is false (it used to read as true), junk is an error
    r = subprocess.run([str(exe), "-scene", "/nonexistent.json", "-headless=false", "-gpu=false"], capture_output=True, text=True)
    assert r.returncode == 1 and "headless=0" in r.stderr
    r = subprocess.run([str(exe), "-headless=maybe"], capture_output=True, text=True)
    assert r.returncode == 2 and "invalid boolean value" in r.stderr


@pytest.mark.gpu
def test_cpp_driver_renders_png(tmp_path):
    import subprocess
    from PIL import Image
    exe = ROOT / "path_trace_golang_b200" / "ptb_render"
    out = tmp_path / "c.png"
    r = subprocess.run([str(exe), "-headless", "-scene", str(scene_path("metal_glass_room")), "-width=192", "-height", "108", "-spp", "8",
                        "-depth", "16", "-seed", "3", "-out", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    img = np.array(Image.open(out))
    assert img.shape == (108, 192, 4) and (img[..., 3] == 255).all() and img[..., :3].std() > 3
    # same seed through the Python veneer gives the same bytes (both go through engine.RenderInto of the C++ host mirror)
    from path_trace_golang_b200 import engine, scene
    ref = engine.Render(scene.Load(scene_path("metal_glass_room")), engine.RenderConfig(192, 108, 8, 16), seed=3)
    assert (img == ref).all()


def _build_c_example(tmp_path):
    import subprocess
    exe = tmp_path / "render_c"
    lib_dir = ROOT / "path_trace_golang_b200"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", f"-I{ROOT / 'include'}", "-o", str(exe), str(ROOT / "examples" / "render_c.c"),
                           f"-L{lib_dir}", "-lptb200", f"-Wl,-rpath,{lib_dir}"])
    return exe


def test_c_example_builds_and_fails_loudly_without_gpu(tmp_path):
    """examples/render_c.c drives the library from plain C99 (what a cgo binding does); on a box without CUDA it must
    exit 1 with the library's message instead of producing an image."""
    import subprocess
    import torch
    exe = _build_c_example(tmp_path)
    r = subprocess.run([str(exe), "/nonexistent.json"], capture_output=True, text=True)
    assert r.returncode == 1 and "open scene" in r.stderr
    if not torch.cuda.is_available():
        out = tmp_path / "o.png"
        r = subprocess.run([str(exe), str(scene_path("example_simple")), str(out), "64", "36", "2", "4"], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stderr and not out.exists()


@pytest.mark.gpu
def test_c_example_renders(tmp_path):
    import subprocess
    from PIL import Image
    exe = _build_c_example(tmp_path)
    out = tmp_path / "o.png"
    r = subprocess.run([str(exe), str(scene_path("example_simple")), str(out), "160", "90", "16", "8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "progress callbacks" in r.stdout
    img = np.array(Image.open(out))
    assert img.shape == (90, 160, 4) and (img[..., 3] == 255).all() and img[..., :3].std() > 5
