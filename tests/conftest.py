import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than ~20 s on CPU")


def golden(name: str):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.skip(f"golden fixture {name} not generated")
    return np.load(path)


@pytest.fixture(scope="session")
def rt():
    """The product library through its Python binding; GPU tests fail loudly if it is not built."""
    import petershirleyraytracer_b200 as rt_mod
    rt_mod.lib()
    return rt_mod


@pytest.fixture(scope="session")
def book():
    from petershirleyraytracer_b200 import scenes
    return scenes.book_scene(11)


@pytest.fixture(scope="session")
def default_scene():
    from petershirleyraytracer_b200 import scenes
    return scenes.default_scene()
