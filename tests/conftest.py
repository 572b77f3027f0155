import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracer_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # a fresh checkout has no built artefacts (they are git-ignored): build them once, like __graft_entry__.build()
    need = [os.path.join(ROOT, "cuda-raytracer_b200", "libb2rt.so"), os.path.join(ROOT, "oracle", "liboracle.so")]
    if not all(os.path.exists(f) for f in need):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def root():
    return ROOT


def scene_path(name):
    return os.path.join(ROOT, "scenes", name + ".b2s")
