import json
import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

SCENES = ["example_simple", "test_scene", "metal_glass_room", "test_comprehensive", "gpu_showcase"]
# scene -> (W, H, spp, depth)
CONFIGS = {
    "C1": ("example_simple", 640, 360, 16, 8),
    "C2": ("test_scene", 1920, 1080, 64, 10),
    "C3": ("metal_glass_room", 3840, 2160, 256, 16),
    "C5": ("gpu_showcase", 7680, 4320, 1024, 12),
}
SCENE_DEPTH = {"example_simple": 8, "test_scene": 10, "metal_glass_room": 16, "test_comprehensive": 10, "gpu_showcase": 12}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def scene_path(name):
    return ROOT / "scenes" / f"{name}.json"


def scene_json(name):
    with open(scene_path(name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def oracle_scenes(oracle_mod):
    return {n: oracle_mod.OracleScene.load(scene_path(n)) for n in SCENES}


@pytest.fixture(scope="session")
def ctx():
    """One CUDA context for the whole GPU test session (fails loudly if the library or the GPU is missing)."""
    from path_trace_golang_b200 import engine
    c = engine.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def host_scenes():
    from path_trace_golang_b200 import scene
    return {n: scene.Load(scene_path(n)) for n in SCENES}
