import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def weights():
    return load_golden("weights_calibrated.npz")


@pytest.fixture(scope="session")
def case_a():
    return load_golden("case_a.npz")


@pytest.fixture(scope="session")
def case_b():
    return load_golden("case_b.npz")


@pytest.fixture(scope="session")
def case_bwd():
    return load_golden("case_bwd.npz")


@pytest.fixture(scope="session", autouse=True)
def _native_library():
    """Build libmvsnet_b200.so and the oracle if they are missing (nvcc/gcc cross-compile without a GPU).
    On the GPU box the prebuilt files travel with the snapshot and this is a no-op."""
    from scene_3dreconstruction_mvsnet_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build_library()
    from oracle import oracle as orc
    orc.build()
    try:
        import torch
        # strict fp32 parity: keep cuDNN/cuBLAS off TF32 for the parts that still run on them (FeatureNet)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    except Exception:
        pass
    yield
