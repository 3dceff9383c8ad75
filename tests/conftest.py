import os
import sys

import pytest

os.environ.setdefault("OMP_NUM_THREADS", "4")   # the oracle's XNNPACK shim uses OpenMP; keep it from fanning out over 200 cores

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ORACLE_LIB = os.path.join(ROOT, "oracle", "_ref", "liboracle_ref.so")
ENGINE_LIB = os.path.join(ROOT, "onnxstream_b200", "csrc", "libonnxstream_b200.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle_lib():
    if not os.path.exists(ORACLE_LIB):
        pytest.skip("oracle/_ref/liboracle_ref.so not built (needs /root/reference; run __graft_entry__.build())")
    return ORACLE_LIB


@pytest.fixture(scope="session")
def engine_lib():
    assert os.path.exists(ENGINE_LIB), "libonnxstream_b200.so not built: run __graft_entry__.build()"
    return ENGINE_LIB
