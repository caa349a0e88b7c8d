import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_outputs.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden():
    return np.load(GOLDEN)


@pytest.fixture(scope="session")
def lib_built():
    """Make sure the shared library exists (cross-compiles on a CPU-only box)."""
    from embeddingnet_b200 import build

    return build.build()


def unit_rows(x):
    ss = np.sum(x.astype(np.float64) ** 2, axis=1, keepdims=True)
    return (x / np.sqrt(np.maximum(ss, 1e-12))).astype(np.float32)


def class_tables(n_classes, per_class, d, normalize):
    """Same construction as tests/golden/make_golden.py."""
    from embeddingnet_b200 import synth

    x, _ = synth.make_numpy(n_classes * per_class, d, n_classes=n_classes, rows_per_class=per_class, noise=0.5,
                            relu=True)
    if normalize:
        x = unit_rows(x)
    return [x[c * per_class:(c + 1) * per_class] for c in range(n_classes)]


MINING_CASES = {
    "ref_cfg": (30, 9, 256, 20, 3, 0.5, True),
    "c1": (40, 12, 128, 32, 8, 0.5, True),
    "small_raw": (8, 6, 32, 5, 4, 0.5, False),
}
