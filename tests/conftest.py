import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
HURR = "1x_HurrDeblur_SubCompact_nf24-nc8_244k_net_g"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def model_dir():
    from upscale_video_b200 import ncnn_model
    return ncnn_model.packaged_model_dir()


@pytest.fixture(scope="session")
def oracle_models(model_dir):
    """Oracle-side (independent reader) copies of the packaged models."""
    from oracle import oracle
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = oracle.read_model(model_dir, name)
        return cache[name]

    return get


def golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name + ".npz"))
