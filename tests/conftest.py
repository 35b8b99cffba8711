import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "wgpu-3dgs-viewer_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def ob():
    """The CPU oracle binding (test infrastructure)."""
    from oracle import binding
    binding.build()
    return binding


@pytest.fixture(scope="session")
def sb():
    """The product's Python harness over the C ABI; fails loudly if the .so is missing."""
    import splat_b200
    if not os.path.exists(splat_b200.lib_path()):
        splat_b200.build()  # fresh checkout: the .so is git-ignored (nvcc cross-compiles without a GPU)
    splat_b200.load()
    return splat_b200


@pytest.fixture(scope="session")
def ctx(sb):
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    c = sb.Context(0)
    yield c
    c.close()


def canonicalise(keys_u32: np.ndarray, idx: np.ndarray):
    """SURVEY §8(c) parity rule (3): order inside an equal-key run is a race in the reference;
    compare after sorting each run by ascending index."""
    order = np.lexsort((idx, keys_u32))
    return keys_u32[order], idx[order]
