import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import oracle_py
    path = os.path.join(ROOT, "oracle", "libcsg_oracle.so")
    if not os.path.exists(path):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return oracle_py.Oracle()


@pytest.fixture(scope="session")
def ref_cpu():
    import oracle_py
    if not oracle_py.have_ref_cpu():
        pytest.skip("oracle/_ref/libref_cpu.so not built (needs /root/reference)")
    return oracle_py.RefCPU()


@pytest.fixture(scope="session")
def ref_gpu():
    import oracle_py
    if not oracle_py.have_ref_gpu():
        pytest.skip("oracle/_ref/libref_gpu.so not built (needs /root/reference)")
    return oracle_py.RefGPU()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "ref_gpu_golden.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/ref_gpu_golden.npz not generated yet")
    return np.load(path)


@pytest.fixture(scope="session")
def csg():
    import csg_b200
    return csg_b200
