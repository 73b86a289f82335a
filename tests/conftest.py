import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import tslb200_loader  # noqa: E402

tslb200_loader.load_package()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by the driver with -m gpu)")


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without a GPU skips the gpu-marked tests instead of failing them.  With a GPU
    nothing is skipped: a missing library then fails loudly (tsl_sdr_b200._lib has no fallback)."""
    if _gpu_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device on this box (the product has no CPU path)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import pyoracle
    pyoracle.build()
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference's own objects (oracle/_ref); skipped where they were never built."""
    import pyoracle
    pyoracle.build()
    if not pyoracle.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return pyoracle.Ref("fma")


@pytest.fixture(scope="session")
def pkg():
    import tsl_sdr_b200
    return tsl_sdr_b200


def rand_iq(n, seed, amp=3000.0):
    rng = np.random.default_rng(seed)
    x = rng.normal(0, amp, 2 * n)
    return np.clip(np.round(x), -32768, 32767).astype(np.int16)
