import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "openmm-atmmetaforce-plugin_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are skipped automatically where no device exists so that a plain `pytest tests` stays green
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def abfe():
    return dict(np.load(os.path.join(GOLDEN, "temoa_g1_abfe.npz")))


@pytest.fixture(scope="session")
def rbfe():
    return dict(np.load(os.path.join(GOLDEN, "temoa_g1_g4_rbfe.npz")))


def pytest_sessionfinish(session, exitstatus):
    """GPU runs leave the measured parity errors behind (gpurun_out/ travels back from the GPU box)."""
    try:
        import json
        from helpers import PARITY_LOG
        if PARITY_LOG:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "parity_errors.json"), "w") as fh:
                json.dump(PARITY_LOG, fh, indent=1, sort_keys=True)
    except Exception:
        pass
