import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


def free_port() -> int:
    """A currently unused TCP port on 127.0.0.1 for torchrun rendezvous in multi-process tests."""
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sock:
        sock.bind(("127.0.0.1", 0))
        return sock.getsockname()[1]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    if "kwargs" in d:
        d["kwargs"] = json.loads(str(d["kwargs"]))
    return d


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def ss():
    """The product package with its CUDA library built and loaded (no compute without a GPU)."""
    import splashsurf_b200
    from splashsurf_b200 import build
    build.build()
    splashsurf_b200.load_library()
    if os.environ.get("SS_TEST_EMULATED"):
        # developer switch (never set by the driver): run GPU-marked tests on the CPU executor of the CUDA sources, e.g.
        #   SS_TEST_EMULATED=1 pytest tests -m gpu -k "not full_size and not dam_ and not multi_gpu and not global_big"
        import ctypes
        from test_emulated_pipeline import build_emulated_library
        splashsurf_b200._LIB = splashsurf_b200._bind(ctypes.CDLL(build_emulated_library()))
    return splashsurf_b200


MESH_CASES = ["cfg1_ref", "cube16_ref", "cube16_scalar_ref", "splash_small_ref", "splash_aabb_ref", "global_cube_ref",
              "global_autodisable_ref"]
