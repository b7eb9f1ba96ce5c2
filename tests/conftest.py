import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `pytest -m gpu`)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def weights():
    from seeme_b200 import synthetic as S
    return {"denoiser": S.denoiser_state(0), "vae": S.vae_state(0), "pointnet": S.pointnet_state(0),
            "output_scene": S.output_scene_state(0)}


@pytest.fixture(scope="session")
def golden_stages():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "stages.npz")))


@pytest.fixture(scope="session")
def smpl_buffers():
    from seeme_b200 import synthetic as S
    return S.smpl_buffers()
